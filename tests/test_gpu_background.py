"""GPU: k-mer counting for the background models on the device (k_kmer_hist) -- exact counts against
numpy over the sequences as stored, and NullModel / mono-nucleotide tables bit-identical to the host path that is pinned
against the reference (NullModel.cpp:88-326, backgroundDistribution.cpp:50-95)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _set(kind, seed):
    from xxmotif_b200 import hostprep
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        n, codes = hostprep.synthetic_set(4000, 200, seed=seed)
    elif kind == "ragged_n":
        n = rng.integers(1, 300, size=3000).astype(np.int32)
        codes = rng.integers(1, 5, size=int(n.sum())).astype(np.uint8)
        for _ in range(2000):
            p = int(rng.integers(0, codes.size - 8))
            codes[p:p + int(rng.integers(1, 8))] = 0
    else:  # tiny sequences around the k-mer length and word boundaries
        n = rng.choice(np.array([1, 2, 3, 8, 9, 10, 31, 32, 33, 63, 64, 65], dtype=np.int32), size=2000)
        codes = rng.integers(1, 5, size=int(n.sum())).astype(np.uint8)
        codes[rng.random(codes.size) < 0.03] = 0
    return n, codes


@pytest.mark.parametrize("kind,revcomp", [("uniform", False), ("ragged_n", False), ("ragged_n", True), ("tiny", True), ("tiny", False)])
def test_device_kmer_counts_and_models(kind, revcomp):
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    n, codes = _set(kind, 7)
    n2, c2 = hostprep.revcomp_double(n, codes) if revcomp else (n, codes)
    ctx = xb.Context(0)
    try:
        ctx.load_seqset(n, codes, revcomp=revcomp)
        for order in (0, 1, 2, 4, 5, 7, 8):  # 7: two passes over the stream, 8: six
            got, ms = hostprep.kmer_counts_device(ctx, order)
            want = hostprep.kmer_counts_host(n2, c2, order)
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, "order %d: %d of %d counters differ, first %d: %d vs %d" % (order, bad.size, got.size, bad[0], got[bad[0]], want[bad[0]])
            assert ms > 0
        for order in (2, 4, 8):
            rp, rc = hostprep.null_model_device(ctx, order)
            rp2, rc2 = hostprep.null_model(n2, c2, order)
            assert (rp == rp2).all() and (rc == rc2).all()
        bg, bgl = hostprep.background_device(ctx)
        bg2, bgl2 = hostprep.background(n2, c2)
        assert (bg == bg2).all() and (bgl == bgl2).all()
    finally:
        ctx.close()


def test_device_kmer_counts_full_size():
    """1M x 1 kb: totals of every k-mer length (a checksum of the histogram), marginal consistency, and the time."""
    import torch
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a large GPU")
    nseq, length = 1000000, 1000
    n, codes = hostprep.synthetic_set(nseq, length, seed=4)
    ctx = xb.Context(0)
    try:
        ctx.load_seqset(n, codes, revcomp=False)
        for order in (2, 8):
            c, ms = hostprep.kmer_counts_device(ctx, order)
            off = 1
            for k in range(order + 1):
                size = 5 ** (k + 1)
                assert int(c[off:off + size].sum()) == nseq * (length - k)  # no N: every position with k more bases counts once
                off += size
            assert [int(c[a]) for a in range(1, 5)] == [int((codes == a).sum()) for a in range(1, 5)]
            print("kmer_counts order %d: %.3f ms for %d bases" % (order, ms, nseq * length))
    finally:
        ctx.close()
