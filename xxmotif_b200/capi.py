"""ctypes binding of libxxmotif_gpu.so (include/xxmotif_gpu.h).

This is the thin Python face of the C ABI used by the tests and by bench.py; the product is the shared library.
There is no fallback of any kind: if the library is missing or no CUDA device is present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxxmotif_gpu.so")

SITE_DTYPE = np.dtype([("score", np.float32), ("startPos", np.int32), ("sequence", np.int32)])

LIST_SORTED_SITES, LIST_SORTED_SITES_ELONG, LIST_BEST_STARTPOS = 0, 1, 2
ORDER_AS_IS, ORDER_BY_SCORE, ORDER_BY_STARTPOS = 0, 1, 2

EXPORTS = [
    "xxm_create", "xxm_destroy", "xxm_last_error", "xxm_set_options", "xxm_load_seqset", "xxm_seqset_info", "xxm_set_background",
    "xxm_reinitialize", "xxm_tables_info", "xxm_tables_get", "xxm_window_pvals", "xxm_scan_full", "xxm_rescore", "xxm_scan_oops",
    "xxm_rescore_startposlist", "xxm_update_pwm_with_best_subset", "xxm_startposlist_best_subset", "xxm_startposlist_from_elong",
    "xxm_set_binding_sites", "xxm_get_sites", "xxm_logodds_hist", "xxm_get_stats", "xxm_stream", "xxm_host_tables", "xxm_host_best_subset", "xxm_host_pair_table",
    "xxm_comm_unique_id", "xxm_comm_init", "xxm_comm_destroy", "xxm_optimize_step", "xxm_step_phases", "xxm_device_tables", "xxm_trace_dump", "xxm_elong_keep", "xxm_startposlist_from_kept", "xxm_check_guards", "xxm_set_global_counts", "xxm_seqset_logsum", "xxm_seqset_logsum_from", "xxm_sites_count", "xxm_last_load_info", "xxm_count_gapped_kmers", "xxm_multi_count_gapped_kmers",
    "xxm_multi_create", "xxm_multi_create_on", "xxm_multi_destroy", "xxm_multi_last_error", "xxm_multi_devices", "xxm_multi_set_options", "xxm_multi_set_background",
    "xxm_multi_load_seqset", "xxm_multi_seqset_info", "xxm_multi_reinitialize", "xxm_multi_scan_full", "xxm_multi_rescore", "xxm_multi_scan_oops",
    "xxm_multi_rescore_startposlist", "xxm_multi_get_sites",
]


class XxmOptions(C.Structure):
    _fields_ = [("mops", C.c_int32), ("oops", C.c_int32), ("localization", C.c_int32), ("maxMotifsPerSequence", C.c_int32),
                ("maxSeqCount", C.c_int32), ("plusFrac", C.c_double), ("overrepCorrection", C.c_double), ("bg", C.c_double * 5),
                ("bgLog", C.c_double * 5)]


class XxmStats(C.Structure):
    _fields_ = [("launches", C.c_int64), ("last_scan_ms", C.c_float), ("last_scan_bases", C.c_int64), ("last_scan_hits", C.c_int64),
                ("last_find_ms", C.c_float)]


class XxmStepInfo(C.Structure):
    _fields_ = [("nbSites", C.c_int64), ("nbBestSites", C.c_int64), ("nbSitesGlobal", C.c_int64), ("maxSites", C.c_int64), ("bestMotifNb", C.c_int64),
                ("bestPval", C.c_double), ("updated", C.c_int32)]


class XxmError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libxxmotif_gpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XxmError("libxxmotif_gpu.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C xxmotif_b200/csrc`")
    L = C.CDLL(LIB_PATH)
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    P = C.POINTER
    L.xxm_create.restype = i32
    L.xxm_create.argtypes = [i32, P(vp)]
    L.xxm_destroy.restype = None
    L.xxm_destroy.argtypes = [vp]
    L.xxm_last_error.restype = C.c_char_p
    L.xxm_last_error.argtypes = [vp]
    L.xxm_stream.restype = vp
    L.xxm_stream.argtypes = [vp]
    sig = {
        "xxm_set_options": [vp, P(XxmOptions)],
        "xxm_load_seqset": [vp, i32, vp, vp, i32],
        "xxm_seqset_info": [vp, vp, vp, P(i32), P(i32)],
        "xxm_set_background": [vp, i32, vp, vp],
        "xxm_reinitialize": [vp, vp, vp, i32, i32, i32],
        "xxm_tables_info": [vp, P(i32), P(i32), P(i32), vp, vp, vp],
        "xxm_tables_get": [vp, i32, vp, vp, vp, vp, vp],
        "xxm_window_pvals": [vp, i32, vp, vp, i32, i32, i32, i32, i32, vp, vp],
        "xxm_scan_full": [vp, i32, i32, i32, P(i32), P(i32), P(i32)],
        "xxm_rescore": [vp, i32, i32, i32, i32, i32, P(i32), P(i32)],
        "xxm_scan_oops": [vp, i32, i32, i32, P(i32), P(i32)],
        "xxm_rescore_startposlist": [vp, i32, i32, i32, i32, vp, vp, i32, P(i32), P(i32)],
        "xxm_update_pwm_with_best_subset": [vp, i32, vp, i32, i32, i32, dbl, P(i32), P(i32), P(dbl)],
        "xxm_startposlist_best_subset": [vp, i32, i32, vp, vp, i32, P(i32), P(i32), P(dbl)],
        "xxm_startposlist_from_elong": [vp, i32, i32, vp, vp, i32, P(i32), P(dbl)],
        "xxm_set_binding_sites": [vp, dbl, vp, i32, P(i32)],
        "xxm_get_sites": [vp, i32, i32, i32, vp, i32, P(i32)],
        "xxm_logodds_hist": [vp, vp, vp, i32, i32, vp, vp, i32, P(dbl), P(dbl)],
        "xxm_get_stats": [vp, P(XxmStats)],
        "xxm_host_tables": [vp, vp, i32, i32, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp],
        "xxm_host_best_subset": [vp, vp, i32, i32, i32, i32, i32, i32, i32, dbl, P(i32), P(dbl)],
        "xxm_host_pair_table": [vp, C.c_int64, vp, i32, vp, P(i32)],
        "xxm_comm_unique_id": [vp],
        "xxm_comm_init": [vp, i32, i32, vp],
        "xxm_comm_destroy": [vp],
        "xxm_optimize_step": [vp, vp, vp, i32, i32, i32, i32, i32, dbl, P(XxmStepInfo)],
        "xxm_step_phases": [vp, vp, P(C.c_int64), i32],
        "xxm_trace_dump": [],
        "xxm_check_guards": [vp, P(i32), P(i32)],
        "xxm_set_global_counts": [vp, C.c_int64, C.c_int64, dbl, i32],
        "xxm_seqset_logsum": [vp, P(dbl)],
        "xxm_seqset_logsum_from": [vp, dbl, P(dbl)],
        "xxm_sites_count": [vp, i32, i32, P(i32)],
        "xxm_last_load_info": [vp, P(i32), P(i32)],
        "xxm_count_gapped_kmers": [vp, i32, i32, vp, i32, vp],
        "xxm_elong_keep": [vp],
        "xxm_startposlist_from_kept": [vp, i32, i32, vp, vp, i32, P(i32), P(dbl)],
        "xxm_device_tables": [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, P(C.c_float)],
    }
    # the multi-device layer mirrors the signatures of the calls it wraps
    for name in ("set_options", "set_background", "load_seqset", "seqset_info", "reinitialize", "scan_full", "rescore", "scan_oops", "rescore_startposlist",
                 "get_sites", "count_gapped_kmers"):
        sig["xxm_multi_" + name] = sig["xxm_" + name]
    sig["xxm_multi_create"] = [i32, P(vp)]
    sig["xxm_multi_create_on"] = [vp, i32, P(vp)]
    L.xxm_multi_destroy.restype = None
    L.xxm_multi_destroy.argtypes = [vp]
    L.xxm_multi_last_error.restype = C.c_char_p
    L.xxm_multi_last_error.argtypes = [vp]
    L.xxm_multi_devices.restype = i32
    L.xxm_multi_devices.argtypes = [vp]
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.restype = i32
        fn.argtypes = args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU context == the reference's StartPosUpdater + PVal_Calculator + Kmer_Generator_Reloaded singletons."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        if self.L.xxm_create(device, C.byref(h)):
            raise XxmError(self.L.xxm_last_error(None).decode())
        self.h = h
        self.nent = 0
        self.ncols = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.xxm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise XxmError(self.L.xxm_last_error(self.h).decode())

    # ---- configuration / inputs
    def set_options(self, mops=False, oops=False, localization=False, bg=None, bgLog=None, maxMotifsPerSequence=1000, maxSeqCount=1000,
                    plusFrac=0.0, overrepCorrection=1.05):
        o = XxmOptions()
        o.mops, o.oops, o.localization = int(mops), int(oops), int(localization)
        o.maxMotifsPerSequence, o.maxSeqCount = maxMotifsPerSequence, maxSeqCount
        o.plusFrac, o.overrepCorrection = plusFrac, overrepCorrection
        for i in range(5):
            o.bg[i] = float(bg[i])
            o.bgLog[i] = float(bgLog[i])
        self._ck(self.L.xxm_set_options(self.h, C.byref(o)))

    def load_seqset(self, n, codes, revcomp=False):
        n = np.ascontiguousarray(n, dtype=np.int32)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        assert int(n.sum(dtype=np.int64)) == codes.size
        self._ck(self.L.xxm_load_seqset(self.h, len(n), _p(n), _p(codes), int(revcomp)))
        self.nent = len(n)

    def last_load_info(self):
        """(chunks that went over the bus raw, all chunks) of the last load when both packing paths worked at once, else (0, 0)"""
        a, b = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_last_load_info(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def seqset_info(self):
        total = np.zeros(self.nent + 1, dtype=np.int32)
        avg = np.zeros(self.nent + 1, dtype=np.float64)
        mx, mn = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_seqset_info(self.h, _p(total), _p(avg), C.byref(mx), C.byref(mn)))
        return total, avg, mx.value, mn.value

    def set_background(self, order, rprobs, rconds):
        rp = np.ascontiguousarray(rprobs, dtype=np.float32)
        rc = np.ascontiguousarray(rconds, dtype=np.float32)
        assert rp.size == 1 << (3 * (order + 1)) and rc.size == rp.size
        self._ck(self.L.xxm_set_background(self.h, order, _p(rp), _p(rc)))

    def reinitialize(self, pwm, cols, posSetSize, totalPositions):
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        self._ck(self.L.xxm_reinitialize(self.h, _p(pwm), _p(cols), len(cols), posSetSize, totalPositions))
        self.ncols = len(cols)

    def tables(self):
        Lc, mk, nparts = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        size = np.zeros(2, dtype=np.int64)
        ncls = np.zeros(2, dtype=np.int32)
        nids = np.zeros(2, dtype=np.int64)
        self._ck(self.L.xxm_tables_info(self.h, C.byref(Lc), C.byref(mk), C.byref(nparts), _p(size), _p(ncls), _p(nids)))
        out = {"L": Lc.value, "maxk": mk.value, "nparts": nparts.value, "pvalues": [], "classes": [], "probs": None}
        for part in range(nparts.value):
            pv = np.zeros(size[part], dtype=np.float64)
            pr = np.zeros(size[part], dtype=np.float64) if (part == 0 and nparts.value == 2) else None
            cs = np.zeros(ncls[part], dtype=np.int32)
            cc = np.zeros(ncls[part], dtype=np.int32)
            ci = np.zeros(nids[part], dtype=np.int64)
            self._ck(self.L.xxm_tables_get(self.h, part, _p(pv), _p(pr) if pr is not None else None, _p(cs), _p(cc), _p(ci)))
            out["pvalues"].append(pv)
            out["classes"].append((cs, cc, ci))
            if pr is not None:
                out["probs"] = pr
        return out

    # ---- the hot path
    def window_pvals(self, seq, pos, posSetSize, firstMotifColumn, enr, mops):
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        f = np.ones(len(seq), dtype=np.float64)
        s = np.ones(len(seq), dtype=np.float64)
        self._ck(self.L.xxm_window_pvals(self.h, len(seq), _p(seq), _p(pos), posSetSize, firstMotifColumn, enr[0], enr[1], int(mops), _p(f), _p(s)))
        return f, s

    def scan_full(self, posSetSize, enr):
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_scan_full(self.h, posSetSize, enr[0], enr[1], C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def rescore(self, firstMotifColumn, posSetSize, enr, palin=False):
        a, c = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_rescore(self.h, firstMotifColumn, posSetSize, enr[0], enr[1], int(palin), C.byref(a), C.byref(c)))
        return a.value, c.value

    def scan_oops(self, posSetSize, enr):
        a, c = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_scan_oops(self.h, posSetSize, enr[0], enr[1], C.byref(a), C.byref(c)))
        return a.value, c.value

    def rescore_startposlist(self, posSetSize, enr, siteSeq, sitePos, instanceThreshold=False):
        ss = np.ascontiguousarray(siteSeq, dtype=np.int32)
        sp = np.ascontiguousarray(sitePos, dtype=np.int32)
        a, c = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_rescore_startposlist(self.h, posSetSize, enr[0], enr[1], len(ss), _p(ss), _p(sp), int(instanceThreshold), C.byref(a),
                                                 C.byref(c)))
        return a.value, c.value

    def update_pwm_with_best_subset(self, which, pwm, firstColumn, posSetSize, plus, pseudo):
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155).copy()
        upd, nb, bp = C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        self._ck(self.L.xxm_update_pwm_with_best_subset(self.h, which, _p(pwm), firstColumn, posSetSize, int(plus), pseudo, C.byref(upd), C.byref(nb),
                                                        C.byref(bp)))
        return upd.value, nb.value, bp.value, pwm

    def startposlist_best_subset(self, firstColumn, posSetSize, cap):
        os_ = np.zeros(max(1, cap), dtype=np.int32)
        op = np.zeros(max(1, cap), dtype=np.int32)
        n, nb, bp = C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        self._ck(self.L.xxm_startposlist_best_subset(self.h, firstColumn, posSetSize, _p(os_), _p(op), len(os_), C.byref(n), C.byref(nb), C.byref(bp)))
        k = max(0, n.value)
        return n.value, nb.value, bp.value, os_[:k], op[:k]

    def startposlist_from_elong(self, bestMotifNb, firstTmpColumn):
        os_ = np.zeros(max(1, bestMotifNb), dtype=np.int32)
        op = np.zeros(max(1, bestMotifNb), dtype=np.int32)
        n, w = C.c_int32(0), C.c_double(1.0)
        self._ck(self.L.xxm_startposlist_from_elong(self.h, bestMotifNb, firstTmpColumn, _p(os_), _p(op), len(os_), C.byref(n), C.byref(w)))
        return os_[:n.value], op[:n.value], w.value

    def set_binding_sites(self, thr, cap):
        out = np.zeros(max(1, cap), dtype=SITE_DTYPE)
        n = C.c_int32(0)
        self._ck(self.L.xxm_set_binding_sites(self.h, thr, _p(out), len(out), C.byref(n)))
        return out[:n.value]

    def get_sites(self, which, order=ORDER_AS_IS, cap=None):
        if cap is None:
            cap = 1 << 16
        while True:
            out = np.zeros(max(1, cap), dtype=SITE_DTYPE)
            n = C.c_int32(0)
            rc = self.L.xxm_get_sites(self.h, which, order, 0, _p(out), len(out), C.byref(n))
            if rc and n.value > len(out):
                cap = n.value
                continue
            self._ck(rc)
            return out[:n.value]

    def logodds_hist(self, pwm, cols, posSetSize, npos):
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        fw = np.zeros(max(1, npos), dtype=np.float64)
        bw = np.zeros(max(1, npos), dtype=np.float64)
        m1, m2 = C.c_double(0), C.c_double(0)
        self._ck(self.L.xxm_logodds_hist(self.h, _p(pwm), _p(cols), len(cols), posSetSize, _p(fw), _p(bw), npos, C.byref(m1), C.byref(m2)))
        return fw[:npos], bw[:npos], m1.value, m2.value

    def comm_init(self, rank, nranks, id128):
        """attach an NCCL communicator (one process and one context per GPU); id128 = comm_unique_id() of rank 0"""
        idb = np.ascontiguousarray(id128, dtype=np.uint8)
        assert idb.size == 128
        self._ck(self.L.xxm_comm_init(self.h, int(rank), int(nranks), _p(idb)))

    def optimize_step(self, pwm, cols, posSetSize, totalPositions, enr, pseudo):
        """one optimizeMotif iteration as one native call (xxm_optimize_step): returns (new pwm, info dict)"""
        pw = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155).copy()
        cl = np.ascontiguousarray(cols, dtype=np.int32)
        info = XxmStepInfo()
        self._ck(self.L.xxm_optimize_step(self.h, _p(pw), _p(cl), len(cl), int(posSetSize), int(totalPositions), int(enr[0]), int(enr[1]), float(pseudo),
                                          C.byref(info)))
        return pw, {"nbSites": info.nbSites, "nbBestSites": info.nbBestSites, "nbSites_global": info.nbSitesGlobal, "maxSites": info.maxSites,
                    "bestMotifNb": info.bestMotifNb, "bestPval": info.bestPval, "updated": bool(info.updated)}

    def nval_hint(self):
        """number of value indices of the current model (size of the histogram a step reads back)"""
        L_, mk, npart = C.c_int32(), C.c_int32(), C.c_int32()
        size = (C.c_int64 * 2)()
        ncl = (C.c_int32 * 2)()
        nid = (C.c_int64 * 2)()
        self.L.xxm_tables_info(self.h, C.byref(L_), C.byref(mk), C.byref(npart), size, ncl, nid)
        return int(ncl[0])

    def stats(self):
        s = XxmStats()
        self._ck(self.L.xxm_get_stats(self.h, C.byref(s)))
        return {"launches": s.launches, "last_scan_ms": s.last_scan_ms, "last_scan_bases": s.last_scan_bases, "last_scan_hits": s.last_scan_hits,
                "last_find_ms": s.last_find_ms}

    def device_tables(self, pwm, cols, posSetSize, totalPositions):
        """the p-value table of a one-part model built on the device (xxm_device_tables): the dict Context.tables() returns + device ms"""
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        info = np.zeros(9, dtype=np.int64)
        ms = C.c_float(0)
        self._ck(self.L.xxm_device_tables(self.h, _p(pwm), _p(cols), len(cols), posSetSize, totalPositions, _p(info), None, None, None, None, C.byref(ms)))
        pv = np.zeros(info[3], dtype=np.float64)
        cs = np.zeros(info[5], dtype=np.int32)
        cc = np.zeros(info[5], dtype=np.int32)
        ci = np.zeros(info[7], dtype=np.int64)
        self._ck(self.L.xxm_device_tables(self.h, _p(pwm), _p(cols), len(cols), posSetSize, totalPositions, _p(info), _p(pv), _p(cs), _p(cc), _p(ci), C.byref(ms)))
        return {"L": int(info[0]), "maxk": int(info[1]), "nparts": int(info[2]), "pvalues": [pv], "classes": [(cs, cc, ci)], "probs": None}, ms.value

    def set_global_counts(self, nentGlobal, totalGlobal, sumLogLength=float("nan"), maxLengGlobal=None):
        """counts of the whole sharded set for the split-phase path (xxm_set_global_counts)"""
        if maxLengGlobal is None:
            maxLengGlobal = self.seqset_info()[2]
        self._ck(self.L.xxm_set_global_counts(self.h, int(nentGlobal), int(totalGlobal), float(sumLogLength), int(maxLengGlobal)))

    def seqset_logsum(self):
        v = C.c_double(0.0)
        self._ck(self.L.xxm_seqset_logsum(self.h, C.byref(v)))
        return v.value

    def count_gapped_kmers(self, nseq, offsets, firstHalfOnly=False):
        """ElongCandidates::countAllOccurrences' counting loop: counts[4^k], first character most significant"""
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        out = np.zeros(4 ** len(off), dtype=np.uint32)
        self._ck(self.L.xxm_count_gapped_kmers(self.h, nseq, len(off), _p(off), int(firstHalfOnly), _p(out)))
        return out

    def check_guards(self):
        """(buffers checked, buffers with an overwritten canary) -- needs XXM_DEBUG_GUARDS in the environment at load time"""
        a, b = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.xxm_check_guards(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def step_phases(self, reset=True):
        """average host-side phase times of xxm_optimize_step since the last reset, microseconds per step"""
        us = np.zeros(7, dtype=np.float64)
        k = C.c_int64(0)
        self._ck(self.L.xxm_step_phases(self.h, _p(us), C.byref(k), int(reset)))
        names = ["tables", "launch_scan", "wait_scan", "choose", "launch_count", "wait_count", "pwm"]
        return {nm: (float(v) / k.value if k.value else 0.0) for nm, v in zip(names, us)}, int(k.value)

    def stream(self):
        return self.L.xxm_stream(self.h)


class _MultiEntryPoints:
    """xxm_foo -> xxm_multi_foo, so that MultiContext reuses Context's marshalling; calls without a multi-device form raise."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        if not name.startswith("xxm_") or not hasattr(self._lib, "xxm_multi_" + name[4:]):
            raise XxmError("%s has no multi-device form (include/xxmotif_gpu.h, xxm_multi)" % name)
        return getattr(self._lib, "xxm_multi_" + name[4:])


class MultiContext(Context):
    """xxm_multi: ONE host thread, several GPUs -- the calls the reference-side binding makes, over shards of the set."""

    def __init__(self, ngpus):
        lib = load_library()
        h = C.c_void_p()
        if lib.xxm_multi_create(ngpus, C.byref(h)):
            raise XxmError(lib.xxm_multi_last_error(None).decode())
        self.L = _MultiEntryPoints(lib)
        self._lib = lib
        self.h = h
        self.nent = 0
        self.ncols = 0

    def close(self):
        if getattr(self, "h", None):
            self._lib.xxm_multi_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc:
            raise XxmError(self._lib.xxm_multi_last_error(self.h).decode())


def host_tables(pwm, cols, posSetSize, totalPositions, bgLog, order, rprobs, rconds):
    """Device-free table construction (xxm_host_tables): returns the same dict as Context.tables()."""
    L = load_library()
    pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(155)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    bgl = np.ascontiguousarray(bgLog, dtype=np.float64)
    rp = np.ascontiguousarray(rprobs, dtype=np.float32)
    rc = np.ascontiguousarray(rconds, dtype=np.float32)
    info = np.zeros(9, dtype=np.int64)
    rcode = L.xxm_host_tables(_p(pwm), _p(cols), len(cols), posSetSize, totalPositions, _p(bgl), order, _p(rp), _p(rc), -1, _p(info), None, None, None,
                              None, None)
    if rcode:
        raise XxmError("xxm_host_tables failed with %d" % rcode)
    out = {"L": int(info[0]), "maxk": int(info[1]), "nparts": int(info[2]), "pvalues": [], "classes": [], "probs": None}
    for part in range(out["nparts"]):
        pv = np.zeros(info[3 + part], dtype=np.float64)
        pr = np.zeros(info[3 + part], dtype=np.float64) if (part == 0 and out["nparts"] == 2) else None
        cs = np.zeros(info[5 + part], dtype=np.int32)
        cc = np.zeros(info[5 + part], dtype=np.int32)
        ci = np.zeros(info[7 + part], dtype=np.int64)
        rcode = L.xxm_host_tables(_p(pwm), _p(cols), len(cols), posSetSize, totalPositions, _p(bgl), order, _p(rp), _p(rc), part, _p(info), _p(pv),
                                  _p(pr) if pr is not None else None, _p(cs), _p(cc), _p(ci))
        if rcode:
            raise XxmError("xxm_host_tables failed with %d" % rcode)
        out["pvalues"].append(pv)
        out["classes"].append((cs, cc, ci))
        if pr is not None:
            out["probs"] = pr
    return out


def host_best_subset(score, seq, posSetSize, maxSites, nent, maxSeqCount=1000, maxMotifsPerSequence=1000, oneOccurrence=False, plusFrac=0.0):
    L = load_library()
    score = np.ascontiguousarray(score, dtype=np.float32)
    seq = np.ascontiguousarray(seq, dtype=np.int32)
    nb, bp = C.c_int32(0), C.c_double(0.0)
    rcode = L.xxm_host_best_subset(_p(score), _p(seq), len(score), posSetSize, maxSeqCount, maxSites, nent, maxMotifsPerSequence, int(oneOccurrence),
                                   plusFrac, C.byref(nb), C.byref(bp))
    if rcode:
        raise XxmError("xxm_host_best_subset failed with %d" % rcode)
    return nb.value, bp.value


def host_pair_table(ids, coloff):
    """Device-free pair-table construction (xxm_host_pair_table): (representable, table[1024], levels)."""
    L = load_library()
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    co = np.ascontiguousarray(coloff, dtype=np.int32)
    tab = np.zeros(1024, dtype=np.uint32)
    lv = C.c_int32(0)
    rcode = L.xxm_host_pair_table(_p(ids), len(ids), _p(co), len(co), _p(tab), C.byref(lv))
    if rcode not in (0, 1):
        raise XxmError("xxm_host_pair_table failed with %d" % rcode)
    return rcode == 0, tab, lv.value


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0); distribute the 128 bytes to the other ranks"""
    L = load_library()
    idb = np.zeros(128, dtype=np.uint8)
    if L.xxm_comm_unique_id(_p(idb)):
        raise XxmError(L.xxm_last_error(None).decode())
    return idb
