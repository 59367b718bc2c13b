"""ctypes binding of oracle/liboracle.so (the C restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")

SITE_DTYPE = np.dtype([("score", np.float32), ("startPos", np.int32), ("sequence", np.int32)])


class XoConfig(C.Structure):
    _fields_ = [("revcomp", C.c_int32), ("mops", C.c_int32), ("oops", C.c_int32), ("localization", C.c_int32),
                ("maxMotifsPerSequence", C.c_int32), ("maxSeqCount", C.c_int32), ("plusFrac", C.c_double),
                ("overrepCorrection", C.c_double), ("bg", C.c_double * 5), ("bgLog", C.c_double * 5)]


def _build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ROOT, "oracle", "xxm_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        _build()
    L = C.CDLL(LIB)
    vp, i32, i64, dbl, flt = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float
    P = C.POINTER
    sig = {
        "xo_seqset_new": (vp, [i32, vp, vp]), "xo_seqset_free": (None, [vp]), "xo_seqset_features": (None, [vp]),
        "xo_seqset_revcomp": (None, [vp]), "xo_seqset_nent": (i32, [vp]), "xo_seqset_max_leng": (i32, [vp]),
        "xo_seqset_min_leng": (i32, [vp]), "xo_seqset_n": (P(i32), [vp]), "xo_seqset_total": (P(i32), [vp]),
        "xo_seqset_avgLength": (P(dbl), [vp]), "xo_seqset_flat": (i64, [vp, vp]),
        "xo_background": (None, [vp, vp, vp]), "xo_nullmodel_new": (vp, [vp, flt, flt, i32]),
        "xo_nullmodel_from_tables": (vp, [i32, vp, vp]), "xo_nullmodel_free": (None, [vp]),
        "xo_nullmodel_rfields": (i32, [vp]), "xo_nullmodel_rprobs": (P(flt), [vp]), "xo_nullmodel_rconds": (P(flt), [vp]),
        "xo_tables_new": (vp, [vp, vp, i32, i32, i32, vp, vp]), "xo_tables_free": (None, [vp]),
        "xo_tables_error": (i32, [vp]), "xo_tables_L": (i32, [vp]), "xo_tables_maxk": (i32, [vp]),
        "xo_tables_nparts": (i32, [vp]), "xo_tables_size": (i64, [vp, i32]), "xo_tables_pvalues": (P(dbl), [vp, i32]),
        "xo_tables_probs": (P(dbl), [vp]), "xo_tables_nclasses": (i32, [vp, i32]),
        "xo_tables_class_score": (P(i32), [vp, i32]), "xo_tables_class_count": (P(i32), [vp, i32]),
        "xo_tables_class_ids": (P(i64), [vp, i32]), "xo_tables_lookup_window": (dbl, [vp, vp]),
        "xo_ctx_new": (vp, [vp, P(XoConfig)]), "xo_ctx_free": (None, [vp]), "xo_ctx_best": (vp, [vp]),
        "xo_ctx_sites": (vp, [vp]), "xo_ctx_elong": (vp, [vp]), "xo_ctx_nbSites": (i32, [vp]),
        "xo_ctx_nbBestSites": (i32, [vp]), "xo_ctx_maxSites": (i32, [vp]), "xo_ctx_positionalProb": (P(dbl), [vp]),
        "xo_init_positional_probs": (None, [vp, i32, i32, i32, i32, i32]),
        "xo_calculate_pval": (dbl, [vp, vp, i32, i32, i32, i32, i32, dbl, i32, P(dbl)]),
        "xo_scan_full": (None, [vp, vp, vp, i32, i32, i32, i32, i32]),
        "xo_update_pwm_with_best_subset": (i32, [vp, i32, vp, vp, i32, i32, i32, i32, dbl, P(i32), P(dbl)]),
        "xo_rescore": (None, [vp, vp, i32, vp, i32, i32, i32, i32, i32]),
        "xo_scan_oops": (None, [vp, vp, vp, i32, i32, i32, i32]),
        "xo_set_binding_sites": (i32, [vp, dbl]),
        "xo_rescore_startposlist": (None, [vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, i32]),
        "xo_startposlist_best_subset": (i32, [vp, vp, i32, i32, i32, vp, vp, i32, P(i32), P(dbl)]),
        "xo_startposlist_from_elong": (i32, [vp, i32, i32, vp, vp, i32, P(dbl)]),
        "xo_sort_by_score": (None, [vp, i32]), "xo_sort_by_startpos": (None, [vp, i32, i32]),
        "xo_update_pwm_ool": (None, [vp, vp, vp, i32, i32, i32, dbl]),
        "xo_order_statistics_pvalue": (dbl, [vp, i32, i32, dbl]), "xo_fast_log": (flt, [flt]),
        "xo_logodds_hist": (None, [vp, vp, vp, i32, i32, vp, vp, i32, P(dbl), P(dbl)]),
        "xo_positional_pval": (dbl, [vp, i32, i32, i32, vp, i32]),
        "xo_count_gapped_kmers": (None, [vp, i32, i32, vp, i32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(p, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


class SeqSet:
    """xo_seqset wrapper. `n`/`flat` are PRE-revcomp lengths / concatenated codes (0=N,1=A,2=C,3=G,4=T)."""

    def __init__(self, n, flat, revcomp=False, recount=False):
        self.L = lib()
        n = np.ascontiguousarray(n, dtype=np.int32)
        flat = np.ascontiguousarray(flat, dtype=np.uint8)
        assert int(n.sum()) == flat.size
        self.h = self.L.xo_seqset_new(len(n), _ptr(n), _ptr(flat))
        if revcomp:
            self.L.xo_seqset_revcomp(self.h)
        if recount:  # negSet order of operations, Globals.cpp:491-501
            self.L.xo_seqset_features(self.h)

    def __del__(self):
        try:
            self.L.xo_seqset_free(self.h)
        except Exception:
            pass

    @property
    def nent(self):
        return self.L.xo_seqset_nent(self.h)

    @property
    def max_leng(self):
        return self.L.xo_seqset_max_leng(self.h)

    @property
    def min_leng(self):
        return self.L.xo_seqset_min_leng(self.h)

    @property
    def n(self):
        return _arr(self.L.xo_seqset_n(self.h), self.nent, np.int32)

    @property
    def total(self):
        return _arr(self.L.xo_seqset_total(self.h), self.nent + 1, np.int32)

    @property
    def avgLength(self):
        return _arr(self.L.xo_seqset_avgLength(self.h), self.nent + 1, np.float64)

    def flat(self):
        sz = self.L.xo_seqset_flat(self.h, None)
        out = np.zeros(sz, dtype=np.uint8)
        self.L.xo_seqset_flat(self.h, _ptr(out))
        return out

    def count_gapped_kmers(self, nseq, offsets, firstHalfOnly=False):
        """ElongCandidates::countAllOccurrences' counting loop (elongationCandidates.cpp:234-262)"""
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        out = np.zeros(4 ** len(off), dtype=np.uint32)
        self.L.xo_count_gapped_kmers(self.h, nseq, len(off), _ptr(off), int(firstHalfOnly), _ptr(out))
        return out

    def background(self):
        bg = np.zeros(5)
        bgl = np.zeros(5)
        self.L.xo_background(self.h, _ptr(bg), _ptr(bgl))
        return bg, bgl


class NullModel:
    def __init__(self, seqset=None, alpha=10.0, beta=0.0, order=2, rprobs=None, rconds=None):
        self.L = lib()
        self.order = order
        if seqset is not None:
            self.h = self.L.xo_nullmodel_new(seqset.h, alpha, beta, order)
        else:
            rp = np.ascontiguousarray(rprobs, dtype=np.float32)
            rc = np.ascontiguousarray(rconds, dtype=np.float32)
            assert rp.size == 1 << (3 * (order + 1))
            self.h = self.L.xo_nullmodel_from_tables(order, _ptr(rp), _ptr(rc))

    def __del__(self):
        try:
            self.L.xo_nullmodel_free(self.h)
        except Exception:
            pass

    @property
    def rfields(self):
        return self.L.xo_nullmodel_rfields(self.h)

    @property
    def rprobs(self):
        return _arr(self.L.xo_nullmodel_rprobs(self.h), self.rfields, np.float32)

    @property
    def rconds(self):
        return _arr(self.L.xo_nullmodel_rconds(self.h), self.rfields, np.float32)


class Tables:
    def __init__(self, pwm, cols, posSetSize, totalPositions, bgLog, nm):
        self.L = lib()
        self.pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(31 * 5)
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)
        bgl = np.ascontiguousarray(bgLog, dtype=np.float64)
        self.nm = nm
        self.h = self.L.xo_tables_new(_ptr(self.pwm), _ptr(self.cols), len(self.cols), posSetSize, totalPositions, _ptr(bgl), nm.h)

    def __del__(self):
        try:
            self.L.xo_tables_free(self.h)
        except Exception:
            pass

    @property
    def error(self):
        return self.L.xo_tables_error(self.h)

    @property
    def nparts(self):
        return self.L.xo_tables_nparts(self.h)

    @property
    def ncols(self):
        return self.L.xo_tables_L(self.h)

    @property
    def maxk(self):
        return self.L.xo_tables_maxk(self.h)

    def pvalues(self, part):
        return _arr(self.L.xo_tables_pvalues(self.h, part), self.L.xo_tables_size(self.h, part), np.float64)

    def probs(self):
        return _arr(self.L.xo_tables_probs(self.h), self.L.xo_tables_size(self.h, 0), np.float64)

    def classes(self, part):
        nc = self.L.xo_tables_nclasses(self.h, part)
        score = _arr(self.L.xo_tables_class_score(self.h, part), nc, np.int32)
        count = _arr(self.L.xo_tables_class_count(self.h, part), nc, np.int32)
        ids = _arr(self.L.xo_tables_class_ids(self.h, part), int(count.sum()), np.int64)
        return score, count, ids

    def lookup_window(self, window):
        w = np.ascontiguousarray(window, dtype=np.uint8)
        return self.L.xo_tables_lookup_window(self.h, _ptr(w))


class Ctx:
    """StartPosUpdater + PVal_Calculator state of the restatement."""

    def __init__(self, seqset, revcomp, mops, oops, localization, bg, bgLog, maxMotifsPerSequence=1000, maxSeqCount=1000,
                 plusFrac=0.0, overrepCorrection=1.05):
        self.L = lib()
        self.seqset = seqset
        cfg = XoConfig()
        cfg.revcomp, cfg.mops, cfg.oops, cfg.localization = int(revcomp), int(mops), int(oops), int(localization)
        cfg.maxMotifsPerSequence, cfg.maxSeqCount = maxMotifsPerSequence, maxSeqCount
        cfg.plusFrac, cfg.overrepCorrection = plusFrac, overrepCorrection
        for i in range(5):
            cfg.bg[i] = float(bg[i])
            cfg.bgLog[i] = float(bgLog[i])
        self.cfg = cfg
        self.h = self.L.xo_ctx_new(seqset.h, C.byref(cfg))

    def __del__(self):
        try:
            self.L.xo_ctx_free(self.h)
        except Exception:
            pass

    def _sites(self, p, n):
        if n == 0:
            return np.zeros(0, dtype=SITE_DTYPE)
        buf = (C.c_char * (n * SITE_DTYPE.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=SITE_DTYPE, count=n).copy()

    @property
    def nbSites(self):
        return self.L.xo_ctx_nbSites(self.h)

    @property
    def nbBestSites(self):
        return self.L.xo_ctx_nbBestSites(self.h)

    @property
    def maxSites(self):
        return self.L.xo_ctx_maxSites(self.h)

    def best(self, n=None):
        return self._sites(self.L.xo_ctx_best(self.h), self.nbBestSites if n is None else n)

    def sites(self, n=None):
        return self._sites(self.L.xo_ctx_sites(self.h), self.nbSites if n is None else n)

    def elong(self, n=None):
        return self._sites(self.L.xo_ctx_elong(self.h), self.nbSites if n is None else n)

    def calculate_pval(self, tables, seq, pos, length, enr, posSetSize, mops):
        n = int(self.seqset.n[seq - 1])
        ln = n - length + 1
        region = enr[1] - enr[0] + 1
        d0 = min(ln, region)
        ps = C.c_double(0)
        f = self.L.xo_calculate_pval(self.h, tables.h, seq, pos, ln, d0, length, float(self.seqset.avgLength[posSetSize]), int(mops), C.byref(ps))
        return f, ps.value

    def init_positional_probs(self, motifLength, firstMotifColumn, enr, offset):
        self.L.xo_init_positional_probs(self.h, motifLength, firstMotifColumn, enr[0], enr[1], offset)

    def scan_full(self, tables, cols, posSetSize, enr, keep=True):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        self.L.xo_scan_full(self.h, tables.h, _ptr(cols), len(cols), posSetSize, enr[0], enr[1], int(keep))

    def update_pwm_with_best_subset(self, which, pwm, cols, posSetSize, oldTotalSites, plus, pseudo):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155).copy()
        nb = C.c_int32(0)
        bp = C.c_double(0.0)
        ok = self.L.xo_update_pwm_with_best_subset(self.h, which, _ptr(pwm), _ptr(cols), len(cols), posSetSize, oldTotalSites, int(plus), pseudo,
                                                   C.byref(nb), C.byref(bp))
        return ok, nb.value, bp.value, pwm

    def rescore(self, tables, firstMotifColumn, tmpCols, posSetSize, enr, palin=False):
        tmpCols = np.ascontiguousarray(tmpCols, dtype=np.int32)
        self.L.xo_rescore(self.h, tables.h, firstMotifColumn, _ptr(tmpCols), len(tmpCols), posSetSize, enr[0], enr[1], int(palin))

    def scan_oops(self, tables, cols, posSetSize, enr):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        self.L.xo_scan_oops(self.h, tables.h, _ptr(cols), len(cols), posSetSize, enr[0], enr[1])

    def set_binding_sites(self, thr=1.0):
        return self.L.xo_set_binding_sites(self.h, thr)

    def rescore_startposlist(self, tables, cols, posSetSize, enr, siteSeq, sitePos, instanceThreshold=False):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        ss = np.ascontiguousarray(siteSeq, dtype=np.int32)
        sp = np.ascontiguousarray(sitePos, dtype=np.int32)
        self.L.xo_rescore_startposlist(self.h, tables.h, _ptr(cols), len(cols), posSetSize, enr[0], enr[1], _ptr(ss), _ptr(sp), len(ss),
                                       int(instanceThreshold))

    def startposlist_best_subset(self, cols, posSetSize, oldTotalSites):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        cap = max(1, self.nbSites)
        os_ = np.zeros(cap, dtype=np.int32)
        op = np.zeros(cap, dtype=np.int32)
        nb = C.c_int32(0)
        bp = C.c_double(0.0)
        n = self.L.xo_startposlist_best_subset(self.h, _ptr(cols), len(cols), posSetSize, oldTotalSites, _ptr(os_), _ptr(op), cap, C.byref(nb),
                                               C.byref(bp))
        return n, nb.value, bp.value, os_[:max(n, 0)], op[:max(n, 0)]

    def logodds_hist(self, pwm, cols, posSetSize):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155)
        npos = max(0, self.seqset.max_leng - len(cols) + 1)
        fw = np.zeros(npos)
        bw = np.zeros(npos)
        m1 = C.c_double(0)
        m2 = C.c_double(0)
        self.L.xo_logodds_hist(self.h, _ptr(pwm), _ptr(cols), len(cols), posSetSize, _ptr(fw), _ptr(bw), npos, C.byref(m1), C.byref(m2))
        return fw, bw, m1.value, m2.value

    def positional_pval(self, enrMax, firstMotifColumn, motifLength, sitePos):
        sp = np.ascontiguousarray(sitePos, dtype=np.int32)
        return self.L.xo_positional_pval(self.h, enrMax, firstMotifColumn, motifLength, _ptr(sp), len(sp))


def sort_by_score(sites):
    a = np.ascontiguousarray(sites, dtype=SITE_DTYPE).copy()
    lib().xo_sort_by_score(_ptr(a), len(a))
    return a


def sort_by_startpos(sites, palin=False):
    a = np.ascontiguousarray(sites, dtype=SITE_DTYPE).copy()
    lib().xo_sort_by_startpos(_ptr(a), len(a), int(palin))
    return a
