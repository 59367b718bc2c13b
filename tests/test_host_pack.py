"""Host-side packing of the sequence set (csrc/host_pack.cpp) against a numpy restatement of the device layout
(xxm_device.cuh; reference representation: Alignment.cpp:149-195, createRevcomp seqset.cpp:128-178)."""
import numpy as np
import pytest

from xxmotif_b200 import capi, hostprep


def layout_numpy(n, codes, revcomp):
    """One 64-bit word of (code - 1) << 2j and one N-mask bit j per stored position; every sequence starts on a word
    boundary and is followed by >= 32 positions flagged N; two guard words end the stream."""
    if revcomp:
        n, codes = hostprep.revcomp_double(n, codes)
    n = np.asarray(n, dtype=np.int64)
    wps = (n + 32 + 31) // 32
    woff = np.concatenate(([0], np.cumsum(wps)))
    flat = np.zeros(int(woff[-1] + 2) * 32, dtype=np.uint8)
    src = 0
    for i, L in enumerate(n.tolist()):
        flat[woff[i] * 32:woff[i] * 32 + L] = codes[src:src + L]
        src += L
    b = flat.reshape(-1, 32).astype(np.uint64)
    isn = b == 0
    val = np.where(isn, np.uint64(0), b - np.uint64(1))
    c64 = (val << (2 * np.arange(32, dtype=np.uint64))).sum(axis=1, dtype=np.uint64)
    m32 = (isn.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return c64, m32


@pytest.mark.parametrize("revcomp", [False, True])
def test_pack_matches_layout(revcomp):
    rng = np.random.default_rng(11)
    lens = [1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 200, 1000, 1001, 7] + rng.integers(1, 300, size=60).tolist()
    n = np.array(lens, dtype=np.int32)
    codes = rng.integers(0, 5, size=int(n.sum()), dtype=np.uint8)
    codes[rng.random(codes.size) < 0.7] |= 1  # fewer N than uniform
    codes = np.minimum(codes, 4).astype(np.uint8)
    c64, m32, notx = hostprep.pack_seqset(n, codes, revcomp)
    e64, e32 = layout_numpy(n, codes, revcomp)
    assert np.array_equal(m32, e32)
    assert np.array_equal(c64, e64)
    off = np.concatenate(([0], np.cumsum(n)))
    assert notx.tolist() == [int((codes[off[i]:off[i + 1]] != 0).sum()) for i in range(len(n))]


def test_pack_all_n_and_runs():
    n = np.array([40, 64, 5], dtype=np.int32)
    codes = np.concatenate([np.zeros(40), np.r_[np.ones(30), np.zeros(4), np.full(30, 4)], np.array([1, 0, 2, 0, 3])]).astype(np.uint8)
    for rc in (False, True):
        c64, m32, notx = hostprep.pack_seqset(n, codes, rc)
        e64, e32 = layout_numpy(n, codes, rc)
        assert np.array_equal(c64, e64) and np.array_equal(m32, e32)
        assert notx.tolist() == [0, 60, 3]


def test_pack_rejects_bad_code():
    n = np.array([50, 70], dtype=np.int32)
    codes = np.ones(120, dtype=np.uint8)
    codes[50 + 37] = 5
    with pytest.raises(capi.XxmError) as e:
        hostprep.pack_seqset(n, codes)
    assert "sequence 2" in str(e.value)
    codes[50 + 37] = 200
    with pytest.raises(capi.XxmError):
        hostprep.pack_seqset(n, codes, True)


@pytest.mark.parametrize("variant", ["XXM_PACK_NO_AVX512", "XXM_PACK_NO_AVX2"])
def test_pack_other_simd_paths_in_subprocess(variant):
    """The widest variant the CPU has (AVX-512 BW, else AVX2, else SSE2) is chosen at load time; XXM_PACK_NO_AVX512 forces the
    AVX2 one, XXM_PACK_NO_AVX2 the SSE2 one, each checked in a fresh process."""
    import os
    import subprocess
    import sys
    if os.environ.get("XXM_PACK_NO_AVX2") or os.environ.get("XXM_PACK_NO_AVX512"):
        pytest.skip("already a forced-variant run")
    env = dict(os.environ, **{variant: "1"})
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_host_pack.py"), "-q", "-x", "-k", "layout or runs or bad or threads"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("revcomp", [False, True])
def test_pack_threads_agree(revcomp):
    """Blocks of sequences packed concurrently (streaming stores into disjoint word ranges, as xxm_load_seqset does chunk by
    chunk) give the words of the single-threaded pass."""
    rng = np.random.default_rng(3)
    n = rng.integers(1, 700, size=5000).astype(np.int32)
    codes = rng.integers(0, 5, size=int(n.sum()), dtype=np.uint8)
    a = hostprep.pack_seqset(n, codes, revcomp, threads=1)
    for t in (2, 5, 8):
        b = hostprep.pack_seqset(n, codes, revcomp, threads=t)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    codes[int(n[:4000].sum()) + 3] = 9
    with pytest.raises(capi.XxmError) as e:
        hostprep.pack_seqset(n, codes, revcomp, threads=4)
    assert "sequence 4001" in str(e.value)
