"""CPU: the callers of the path as the product states them (xxmotif_b200/csrc/refine.cpp: optimizeMotif, maximizeMotifLength, the
per-motif loop of MotifContainer::iterate, and their several-contexts variants xxm_maximize_motif_length_on / xxm_iterate_motif_on /
xxm_iterate_motifs) against the REFERENCE's own runs of those functions (stages A, E, F of the golden traces), with the nine
data-sized C-ABI calls they make served by the oracle (tests/host/oracle_abi_shim.cpp over oracle/liboracle.so) instead of the
GPU library.  Bit-exact, like tests/test_gpu_refine.py on the device.  Test infrastructure only: the product never loads this."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import golden_names
from replay import TraceInputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_shim = None


def _lib(tmp):
    global _shim
    if _shim is not None:
        return _shim
    import oracle_lib as ol
    ol.lib()  # builds oracle/liboracle.so when it is stale
    so = os.path.join(str(tmp), "librefine_oracle.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math",
                           os.path.join(ROOT, "tests", "host", "oracle_abi_shim.cpp"), os.path.join(ROOT, "xxmotif_b200", "csrc", "refine.cpp"),
                           "-o", so, "-L", os.path.join(ROOT, "oracle"), "-l:liboracle.so", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    from xxmotif_b200.refine import XxmMotif, XxmRefineParams
    L = C.CDLL(so)
    vp, i32, P = C.c_void_p, C.c_int32, C.POINTER
    L.shim_wrap.restype = vp
    L.shim_wrap.argtypes = [vp, vp, vp, vp, i32]
    L.shim_free.argtypes = [vp]
    L.shim_calls.restype = C.c_long
    L.shim_calls.argtypes = [vp]
    L.xxm_last_error.restype = C.c_char_p
    L.xxm_last_error.argtypes = [vp]
    L.xxm_optimize_motif.argtypes = [vp, P(XxmMotif), P(XxmRefineParams)]
    L.xxm_update_enriched_region.argtypes = [vp, i32, i32, vp]
    L.xxm_maximize_motif_length_on.argtypes = [P(vp), i32, P(XxmMotif), P(XxmRefineParams), i32, P(i32)]
    L.xxm_iterate_motif_on.argtypes = [P(vp), i32, P(XxmMotif), P(XxmRefineParams), P(i32)]
    L.xxm_iterate_motifs.argtypes = [P(vp), i32, P(XxmMotif), i32, P(XxmRefineParams), i32, P(i32), P(i32), P(i32)]
    _shim = L
    return L


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    return _lib(tmp_path_factory.mktemp("refine_oracle"))


class _Contexts:
    """an oracle context per "GPU context" on one sequence set, wrapped for the shim"""

    def __init__(self, L, ti, n):
        import oracle_lib as ol
        from backends import OracleBackend
        t = ti.t
        self.L = L
        self.ob = OracleBackend(ti, from_scratch="nm.rprobs" not in t)
        self.octx = [self.ob.ctx]
        for _ in range(n - 1):
            self.octx.append(ol.Ctx(self.ob.seqset, ti.revcomp, ti.mops, ti.oops, ti.localization, self.ob.bg, self.ob.bgLog,
                                    maxMotifsPerSequence=int(t["cfg.maxMotifsPerSequence"][0]), maxSeqCount=int(t["cfg.maxSeqCount"][0]),
                                    plusFrac=float(t["cfg.plusFrac"][0]), overrepCorrection=float(t["cfg.overrepCorrection"][0])))
        self.bgLog = np.ascontiguousarray(self.ob.bgLog, dtype=np.float64)
        self.h = [L.shim_wrap(o.h, self.ob.seqset.h, self.ob.nm.h, self.bgLog.ctypes.data_as(C.c_void_p), int(ti.revcomp)) for o in self.octx]

    def arr(self, k):
        return (C.c_void_p * k)(*self.h[:k])

    def close(self):
        for h in self.h:
            self.L.shim_free(h)
        self.h = []


def _motif(t, pss):
    from xxmotif_b200 import refine
    return refine.Motif(t["m0.pwm"], t["m0.cols"], pss, enr=t["m0.enrich"], siteSeq=t["m0.sites.seq"], sitePos=t["m0.sites.pos"],
                        pval=float(t["m0.pval"][0]), posPval=float(t["m0.posPval"][0]))


def _params(ti, pss):
    from xxmotif_b200.refine import XxmRefineParams
    t = ti.t
    return XxmRefineParams(int(ti.total[pss]), int(ti.revcomp), int(ti.localization), int(t["cfg.maximizeMotifLength"][0]), int(t["cfg.minMatchPositions"][0]),
                           int(t["cfg.maxMatchPositions"][0]), float(ti.pseudo), float(t["cfg.neff_pwm"][0]), None)


def _run_stages(L, ctxs, nctx, ti):
    """stages A, E, F of ref_harness on ctxs.h[0] (+ nctx - 1 more contexts for the column variants), each against the trace"""
    from test_gpu_refine import _cmp_motif
    t = ti.t
    pss = int(t["m0.posSetSize"][0])
    m = _motif(t, pss)
    p = _params(ti, pss)
    elong = np.zeros(31 * 5, dtype=np.float64)  # IterateMotif::elongPWM
    p.elongPWM = elong.ctypes.data
    owner = ctxs.h[0]
    assert L.xxm_optimize_motif(owner, C.byref(m.m), C.byref(p)) == 0, L.xxm_last_error(owner)
    assert np.array_equal(m.pwm[5:], t["opt1.pwm_out"][5:])
    assert m.posPval == float(t["optA.posPval"][0])
    if ti.localization:
        cols = m.cols
        e = np.array(m.enr, dtype=np.int32)
        assert L.xxm_update_enriched_region(owner, int(cols[0]), int(cols[-1] - cols[0] + 1), e.ctypes.data_as(C.c_void_p)) == 0
        for i in range(3):
            m.m.enr[i] = int(e[i])
        assert np.array_equal(m.enr, t["optA.enriched.enrich"])
    nc = C.c_int32(0)
    assert L.xxm_maximize_motif_length_on(ctxs.arr(nctx), nctx, C.byref(m.m), C.byref(p), int(t["cfg.maximizeMotifLength"][0]), C.byref(nc)) == 0
    _cmp_motif("maxlen on %d contexts" % nctx, m, t, "maxlen")
    r = C.c_int32(0)
    assert L.xxm_iterate_motif_on(ctxs.arr(nctx), nctx, C.byref(m.m), C.byref(p), C.byref(r)) == 0
    assert r.value >= 1
    _cmp_motif("converged on %d contexts" % nctx, m, t, "converged")
    return m


@pytest.mark.parametrize("name", golden_names())
def test_refine_callers_against_reference_on_cpu(shim, golden, name):
    t = golden(name)
    ti = TraceInputs(t)
    ctxs = _Contexts(shim, ti, 3)
    try:
        one = _run_stages(shim, ctxs, 1, ti)
        used = [shim.shim_calls(h) for h in ctxs.h]
        assert used[0] > 0 and used[1] == 0 and used[2] == 0
        three = _run_stages(shim, ctxs, 3, ti)
        if int(t["cfg.maximizeMotifLength"][0]):
            assert all(shim.shim_calls(h) > 0 for h in ctxs.h)  # the variants did spread
        assert one.pwm.tobytes() == three.pwm.tobytes() and one.pval == three.pval
    finally:
        ctxs.close()


@pytest.mark.parametrize("name", ["c1_zoops", "c2_mops_rc", "c5_loc_gap"])
def test_iterate_motifs_on_cpu(shim, golden, name):
    """xxm_iterate_motifs: the trace's motif and shifted / trimmed copies of it, one after the other on one context and three at a
    time; and the trace's own motif still ends where the reference's per-motif loop ends"""
    from xxmotif_b200 import refine
    from xxmotif_b200.refine import XxmMotif
    t = golden(name)
    ti = TraceInputs(t)
    pss = int(t["m0.posSetSize"][0])
    ctxs = _Contexts(shim, ti, 3)
    try:
        def motifs():
            ms = [_motif(t, pss)]
            cols = [int(c) for c in t["m0.cols"]]
            for cs in (cols[:-1], cols[1:], cols[:-2]):  # sub-motifs of the start model: valid, normalised columns
                if len(cs) >= 5:
                    ms.append(refine.Motif(t["m0.pwm"], cs, pss, enr=t["m0.enrich"], siteSeq=t["m0.sites.seq"], sitePos=t["m0.sites.pos"],
                                           pval=float(t["m0.pval"][0]), posPval=float(t["m0.posPval"][0])))
            return ms

        def run(nctx):
            ms = motifs()
            arr = (XxmMotif * len(ms))()
            for i, m in enumerate(ms):
                C.memmove(C.byref(arr[i]), C.byref(m.m), C.sizeof(XxmMotif))
            p = _params(ti, pss)
            rounds = (C.c_int32 * len(ms))()
            bad, badc = C.c_int32(-1), C.c_int32(-1)
            rc = shim.xxm_iterate_motifs(ctxs.arr(nctx), nctx, arr, len(ms), C.byref(p), 1, rounds, C.byref(bad), C.byref(badc))
            assert rc == 0, (rc, bad.value)
            for i, m in enumerate(ms):
                C.memmove(C.byref(m.m), C.byref(arr[i]), C.sizeof(XxmMotif))
            return [(m.cols.tolist(), m.pwm.tobytes(), m.enr.tolist(), m.pval, m.posPval, m.siteSeq.tobytes(), m.sitePos.tobytes()) for m in ms], list(rounds)

        s1, r1 = run(1)
        s3, r3 = run(3)
        assert len(s1) >= 2 and s1 == s3 and r1 == r3
    finally:
        ctxs.close()
