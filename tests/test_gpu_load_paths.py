"""GPU: the three packing paths of xxm_load_seqset (host cores, csrc/host_pack.cpp, when XXM_PACK_THREADS >= 6; k_seq_stats +
k_encode on the device with a single thread; both at once -- raw chunks from the front, host-packed chunks from the back -- in
between) leave the same resident set: exact k-mer counts of the stored stream (forward and
doubled) against the host count over the reference's representation, and identical total[] / avgLength[]."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("revcomp", [False, True])
def test_host_and_device_packing_agree(revcomp):
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    rng = np.random.default_rng(23)
    n = rng.integers(1, 2000, size=6000).astype(np.int32)  # ~6 MB of bases: above the single-thread shortcut
    n[:64] = np.arange(1, 65)                               # every length around the word boundaries
    codes = rng.integers(1, 5, size=int(n.sum())).astype(np.uint8)
    for _ in range(3000):
        p = int(rng.integers(0, codes.size - 40))
        codes[p:p + int(rng.integers(1, 40))] = 0
    n2, c2 = hostprep.revcomp_double(n, codes) if revcomp else (n, codes)
    want = {order: hostprep.kmer_counts_host(n2, c2, order) for order in (2, 5)}
    saved = os.environ.get("XXM_PACK_THREADS")
    infos = []
    try:
        for threads, path in (("8", "host"), ("1", "device"), ("3", "hybrid")):
            os.environ["XXM_PACK_THREADS"] = threads
            os.environ["XXM_LOAD_PATH"] = path
            ctx = xb.Context(0)
            try:
                ctx.load_seqset(n, codes, revcomp=revcomp)
                total, avg, mxl, mnl = ctx.seqset_info()
                infos.append((np.array(total), np.array(avg), mxl, mnl))
                for order, w in want.items():
                    got, _ = hostprep.kmer_counts_device(ctx, order)
                    assert np.array_equal(got, w), (threads, order)
                ctx.load_seqset(n, codes, revcomp=revcomp)  # staging buffers are reused by a second load
                got, _ = hostprep.kmer_counts_device(ctx, 2)
                assert np.array_equal(got, want[2]), (threads, "reload")
            finally:
                ctx.close()
    finally:
        os.environ.pop("XXM_LOAD_PATH", None)
        if saved is None:
            os.environ.pop("XXM_PACK_THREADS", None)
        else:
            os.environ["XXM_PACK_THREADS"] = saved
    for other in infos[1:]:
        assert np.array_equal(infos[0][0], other[0]) and np.array_equal(infos[0][1], other[1]) and infos[0][2:] == other[2:]


def test_bad_code_is_reported_by_both_paths():
    import xxmotif_b200 as xb
    rng = np.random.default_rng(5)
    n = np.full(5000, 1000, dtype=np.int32)
    codes = rng.integers(1, 5, size=int(n.sum())).astype(np.uint8)
    codes[3217 * 1000 + 417] = 7
    saved = os.environ.get("XXM_PACK_THREADS")
    try:
        for threads, path in (("8", "host"), ("1", "device"), ("3", "hybrid")):
            os.environ["XXM_PACK_THREADS"] = threads
            os.environ["XXM_LOAD_PATH"] = path
            ctx = xb.Context(0)
            try:
                with pytest.raises(xb.XxmError) as e:
                    ctx.load_seqset(n, codes, revcomp=False)
                assert "3218" in str(e.value)
            finally:
                ctx.close()
    finally:
        os.environ.pop("XXM_LOAD_PATH", None)
        if saved is None:
            os.environ.pop("XXM_PACK_THREADS", None)
        else:
            os.environ["XXM_PACK_THREADS"] = saved


def test_hybrid_load_at_size_leaves_the_host_packed_set():
    """80 MB of bases (several chunks on either side of the meeting point), forced paths: the packed stream read back through the
    k-mer counts, totals and a scan must not depend on who packed which chunk."""
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    n, codes = hostprep.synthetic_set(80000, 1000)
    codes = codes.copy()
    codes[::9973] = 0
    saved = {k: os.environ.get(k) for k in ("XXM_PACK_THREADS", "XXM_LOAD_PATH")}
    out = []
    try:
        os.environ["XXM_PACK_THREADS"] = "3"
        for path in ("host", "hybrid", "device"):
            os.environ["XXM_LOAD_PATH"] = path
            ctx = xb.Context(0)
            try:
                ctx.load_seqset(n, codes, revcomp=True)
                total, avg, mxl, mnl = ctx.seqset_info()
                c5, _ = hostprep.kmer_counts_device(ctx, 5)
                g = ctx.count_gapped_kmers(len(n), [0, 1, 3, 4, 7], 0)
                out.append((np.array(total), np.array(avg), c5, g))
            finally:
                ctx.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for o in out[1:]:
        for a, b in zip(out[0], o):
            assert np.array_equal(a, b)
