"""GPU: the two packing paths of xxm_load_seqset (host cores, csrc/host_pack.cpp, when XXM_PACK_THREADS >= 6; k_seq_stats +
k_encode on the device otherwise) leave the same resident set: exact k-mer counts of the stored stream (forward and
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
        for threads in ("8", "2"):  # host cores, device
            os.environ["XXM_PACK_THREADS"] = threads
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
        if saved is None:
            os.environ.pop("XXM_PACK_THREADS", None)
        else:
            os.environ["XXM_PACK_THREADS"] = saved
    assert np.array_equal(infos[0][0], infos[1][0]) and np.array_equal(infos[0][1], infos[1][1]) and infos[0][2:] == infos[1][2:]


def test_bad_code_is_reported_by_both_paths():
    import xxmotif_b200 as xb
    rng = np.random.default_rng(5)
    n = np.full(5000, 1000, dtype=np.int32)
    codes = rng.integers(1, 5, size=int(n.sum())).astype(np.uint8)
    codes[3217 * 1000 + 417] = 7
    saved = os.environ.get("XXM_PACK_THREADS")
    try:
        for threads in ("8", "2"):
            os.environ["XXM_PACK_THREADS"] = threads
            ctx = xb.Context(0)
            try:
                with pytest.raises(xb.XxmError) as e:
                    ctx.load_seqset(n, codes, revcomp=False)
                assert "3218" in str(e.value)
            finally:
                ctx.close()
    finally:
        if saved is None:
            os.environ.pop("XXM_PACK_THREADS", None)
        else:
            os.environ["XXM_PACK_THREADS"] = saved
