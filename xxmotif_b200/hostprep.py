"""Host-side input preparation for the hot path: reverse-complement doubling, mono-nucleotide background and the
order-k NullModel tables (device-free entry points of the C ABI), plus synthetic planted-motif sequence sets.

Used by bench.py, the tests and the refinement driver; never touches oracle/.
"""
import ctypes as C

import numpy as np

from . import capi

EXTRA_EXPORTS = ["xxm_host_null_model", "xxm_host_background", "xxm_kmer_counts", "xxm_kmer_counts_ms", "xxm_null_model_from_counts",
                 "xxm_null_model_from_seqset", "xxm_background_from_seqset", "xxm_host_pack_seqset"]


def _lib():
    L = capi.load_library()
    if not getattr(L, "_hostprep_ready", False):
        vp, i32, flt = C.c_void_p, C.c_int32, C.c_float
        L.xxm_host_null_model.restype = i32
        L.xxm_host_null_model.argtypes = [i32, vp, vp, i32, flt, flt, vp, vp]
        L.xxm_host_background.restype = i32
        L.xxm_host_background.argtypes = [i32, vp, vp, vp, vp]
        L.xxm_host_pack_seqset.restype = i32
        L.xxm_host_pack_seqset.argtypes = [i32, vp, vp, i32, i32, vp, vp, vp]
        for name, args in {"xxm_kmer_counts": [vp, i32, vp, C.c_int64], "xxm_kmer_counts_ms": [vp, vp],
                           "xxm_null_model_from_counts": [vp, i32, flt, flt, vp, vp], "xxm_null_model_from_seqset": [vp, i32, flt, flt, vp, vp],
                           "xxm_background_from_seqset": [vp, vp, vp]}.items():
            getattr(L, name).restype = i32
            getattr(L, name).argtypes = args
        L._hostprep_ready = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def revcomp_double(n, codes):
    """createRevcomp (seqset.cpp:128-178): every sequence becomes fwd + N + reverse complement. Vectorised."""
    n = np.asarray(n, dtype=np.int64)
    codes = np.asarray(codes, dtype=np.uint8)
    n2 = 2 * n + 1
    out = np.zeros(int(n2.sum()), dtype=np.uint8)
    src_off = np.concatenate(([0], np.cumsum(n)))[:-1]
    dst_off = np.concatenate(([0], np.cumsum(n2)))[:-1]
    comp = np.array([0, 4, 3, 2, 1], dtype=np.uint8)
    if len(set(n.tolist())) == 1:  # uniform length: pure reshape
        L = int(n[0])
        a = codes.reshape(len(n), L)
        o = out.reshape(len(n), 2 * L + 1)
        o[:, :L] = a
        o[:, L + 1:] = comp[a[:, ::-1]]
        return n2.astype(np.int32), out
    for s, d, L in zip(src_off, dst_off, n):
        s, d, L = int(s), int(d), int(L)
        out[d:d + L] = codes[s:s + L]
        out[d + L + 1:d + 2 * L + 1] = comp[codes[s:s + L][::-1]]
    return n2.astype(np.int32), out


def pack_seqset(n, codes, revcomp=False, threads=1):
    """The packing step of xxm_load_seqset alone (xxm_host_pack_seqset): (code words, N masks, residues per sequence).
    Raises on a code outside 0..4."""
    L = _lib()
    n = np.ascontiguousarray(n, dtype=np.int32)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    stored = 2 * n.astype(np.int64) + 1 if revcomp else n.astype(np.int64)
    words = int(((stored + 32 + 31) // 32).sum()) + 2
    c64 = np.zeros(words, dtype=np.uint64)
    m32 = np.zeros(words, dtype=np.uint32)
    notx = np.zeros(len(n), dtype=np.int32)
    bad = L.xxm_host_pack_seqset(len(n), _p(n), _p(codes), int(revcomp), int(threads), _p(c64), _p(m32), _p(notx))
    if bad:
        raise capi.XxmError("sequence %d holds a code outside 0..4" % bad)
    return c64, m32, notx


def background(n, codes):
    """posBg / negBg and logs over the sequences as stored (after revcomp doubling)."""
    n = np.ascontiguousarray(n, dtype=np.int32)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    bg = np.zeros(5)
    bgl = np.zeros(5)
    if _lib().xxm_host_background(len(n), _p(n), _p(codes), _p(bg), _p(bgl)):
        raise capi.XxmError("xxm_host_background failed")
    return bg, bgl


def null_model(n, codes, order=2, alpha=10.0, beta=0.0):
    """NullModel::init over the sequences as stored -> (rprobs, rconds) float32 tables of size 2^(3(order+1))."""
    n = np.ascontiguousarray(n, dtype=np.int32)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    size = 1 << (3 * (order + 1))
    rp = np.zeros(size, dtype=np.float32)
    rc = np.zeros(size, dtype=np.float32)
    if _lib().xxm_host_null_model(len(n), _p(n), _p(codes), order, alpha, beta, _p(rp), _p(rc)):
        raise capi.XxmError("xxm_host_null_model failed")
    return rp, rc


def _fields(order):
    return sum(5 ** k for k in range(order + 2))


def kmer_counts_device(ctx, order):
    """NullModel::init's counting loop over the set loaded in ctx, on the GPU (xxm_kmer_counts): uint64[fields] indexed like
    NullModel::_counts.  Returns (counts, device milliseconds)."""
    c = np.zeros(_fields(order), dtype=np.uint64)
    ctx._ck(_lib().xxm_kmer_counts(ctx.h, order, _p(c), c.size))
    ms = C.c_float(0)
    _lib().xxm_kmer_counts_ms(ctx.h, C.byref(ms))
    return c, ms.value


def kmer_counts_ms(ctx):
    """device milliseconds of the last xxm_kmer_counts / xxm_count_gapped_kmers on this context"""
    ms = C.c_float(0)
    _lib().xxm_kmer_counts_ms(ctx.h, C.byref(ms))
    return ms.value


def kmer_counts_host(n, codes, order):
    """the same counts with numpy (test oracle of the device histogram): all (k+1)-mers, k = 0..order, without N, inside a sequence"""
    n = np.asarray(n, dtype=np.int64)
    codes = np.asarray(codes, dtype=np.uint8)
    c = np.zeros(_fields(order), dtype=np.uint64)
    seqid = np.repeat(np.arange(len(n)), n)
    idx = np.zeros(codes.size, dtype=np.int64)
    ok = np.ones(codes.size, dtype=bool)
    for k in range(order + 1):
        m = codes.size - k
        if m <= 0:
            break
        nxt = codes[k:].astype(np.int64)
        ok = ok[:m] & (nxt != 0) & (seqid[:m] == seqid[k:])
        idx = idx[:m] * 5 + nxt
        np.add.at(c, idx[ok], 1) if ok.sum() < 1 << 16 else c.__iadd__(np.bincount(idx[ok], minlength=c.size).astype(np.uint64))
    return c


def null_model_from_counts(counts, order, alpha=10.0, beta=0.0):
    size = 1 << (3 * (order + 1))
    rp = np.zeros(size, dtype=np.float32)
    rc = np.zeros(size, dtype=np.float32)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    if _lib().xxm_null_model_from_counts(_p(counts), order, alpha, beta, _p(rp), _p(rc)):
        raise capi.XxmError("xxm_null_model_from_counts failed")
    return rp, rc


def null_model_device(ctx, order=2, alpha=10.0, beta=0.0):
    """NullModel::init over the set loaded in ctx, k-mer counting on the GPU -> (rprobs, rconds)"""
    size = 1 << (3 * (order + 1))
    rp = np.zeros(size, dtype=np.float32)
    rc = np.zeros(size, dtype=np.float32)
    ctx._ck(_lib().xxm_null_model_from_seqset(ctx.h, order, alpha, beta, _p(rp), _p(rc)))
    return rp, rc


def background_device(ctx):
    bg = np.zeros(5)
    bgl = np.zeros(5)
    ctx._ck(_lib().xxm_background_from_seqset(ctx.h, _p(bg), _p(bgl)))
    return bg, bgl


IUPAC = {"A": [1], "C": [2], "G": [3], "T": [4], "W": [1, 4], "R": [1, 3], "Y": [2, 4], "S": [2, 3], "K": [3, 4], "M": [1, 2], "N": [1, 2, 3, 4], ".": [1, 2, 3, 4]}


def synthetic_set(nseq, length, motif="TATAWAWR", frac=0.5, seed=4, chunk=1 << 16):
    """i.i.d. uniform ACGT sequences of one length with one planted IUPAC instance in a fraction of them
    (SURVEY.md 8(d) cfg4 shape). Returns (n int32[nseq], codes uint8[nseq*length]) with codes 1..4."""
    rng = np.random.default_rng(seed)
    codes = np.empty(nseq * length, dtype=np.uint8)
    m = len(motif)
    for a in range(0, nseq, chunk):
        b = min(nseq, a + chunk)
        blk = rng.integers(1, 5, size=(b - a, length), dtype=np.uint8)
        has = rng.random(b - a) < frac
        pos = rng.integers(0, length - m + 1, size=b - a)
        rows = np.flatnonzero(has)
        for j, ch in enumerate(motif):
            alts = np.array(IUPAC[ch], dtype=np.uint8)
            blk[rows, pos[rows] + j] = alts[rng.integers(0, len(alts), size=len(rows))]
        codes[a * length:b * length] = blk.reshape(-1)
    return np.full(nseq, length, dtype=np.int32), codes


TATA_PROFILE = np.array([  # the README's 4 x 8 TATA profile shape (rows A, C, G, T), values in percent
    [5, 90, 5, 90, 45, 90, 50, 45],
    [5, 3, 5, 3, 5, 3, 5, 5],
    [5, 3, 5, 3, 5, 3, 5, 45],
    [85, 4, 85, 4, 45, 4, 40, 5]], dtype=np.float64)


def profile_pwm(profile=TATA_PROFILE):
    """Motif::InitProfMotif (Motif.cpp:720-799): a 4 x L profile -> the 31 x 5 log PWM centred in 30 columns and its
    match columns. Values are divided by 100 and renormalised per column; other columns are uniform."""
    L = profile.shape[1]
    pwm = np.zeros((31, 5))
    first = (30 - L) // 2 + 1
    pwm[1:, 1:] = np.log(1.0 / 4)
    for i in range(L):
        col = profile[:, i] / 100.0
        tot = 0.0
        for v in col:
            tot += v
        pwm[first + i, 0] = tot
        for a in range(4):
            pwm[first + i, a + 1] = np.log((col[a] + 0.0 / 1000) / (tot + 4 * 0.0 / 1000))
    return pwm.reshape(-1), np.arange(first, first + L, dtype=np.int32)
