"""ctypes face of the callers of the hot path (csrc/refine.cpp): IterateMotif::optimizeMotif, ::maximizeMotifLength and the
per-motif loop of MotifContainer::iterate, re-stated in C++ above the C ABI (include/xxmotif_gpu.h)."""
import ctypes as C

import numpy as np

from . import capi

EXTRA_EXPORTS = ["xxm_optimize_motif", "xxm_maximize_motif_length", "xxm_iterate_motif", "xxm_iterate_motifs", "xxm_share_seqset", "xxm_view_create",
                 "xxm_view_count", "xxm_share_candidates", "xxm_maximize_motif_length_on", "xxm_iterate_motif_on"]


class XxmMotif(C.Structure):
    _fields_ = [("pwm", C.c_double * (31 * 5)), ("ncols", C.c_int32), ("cols", C.c_int32 * 30), ("posSetSize", C.c_int32), ("enr", C.c_int32 * 3),
                ("pval", C.c_double), ("posPval", C.c_double), ("worstScore", C.c_double), ("nsites", C.c_int32), ("siteCap", C.c_int32),
                ("siteSeq", C.c_void_p), ("sitePos", C.c_void_p)]


class XxmRefineParams(C.Structure):
    _fields_ = [("totalPositions", C.c_int32), ("revcomp", C.c_int32), ("localization", C.c_int32), ("maximizeMotifLength", C.c_int32),
                ("minMatchPositions", C.c_int32), ("maxMatchPositions", C.c_int32), ("pseudo", C.c_double), ("neff_pwm", C.c_double),
                ("elongPWM", C.c_void_p)]


def _lib():
    L = capi.load_library()
    if not getattr(L, "_refine_ready", False):
        vp, i32 = C.c_void_p, C.c_int32
        P = C.POINTER
        L.xxm_optimize_motif.restype = i32
        L.xxm_optimize_motif.argtypes = [vp, P(XxmMotif), P(XxmRefineParams)]
        L.xxm_maximize_motif_length.restype = i32
        L.xxm_maximize_motif_length.argtypes = [vp, P(XxmMotif), P(XxmRefineParams), i32, P(i32)]
        L.xxm_iterate_motif.restype = i32
        L.xxm_iterate_motif.argtypes = [vp, P(XxmMotif), P(XxmRefineParams), P(i32)]
        L.xxm_iterate_motifs.restype = i32
        L.xxm_iterate_motifs.argtypes = [P(vp), i32, P(XxmMotif), i32, P(XxmRefineParams), i32, P(i32), P(i32), P(i32)]
        L.xxm_share_seqset.restype = i32
        L.xxm_share_seqset.argtypes = [vp, vp]
        L.xxm_view_create.restype = i32
        L.xxm_view_create.argtypes = [vp, P(vp)]
        L.xxm_view_count.restype = i32
        L.xxm_view_count.argtypes = [vp]
        L.xxm_share_candidates.restype = i32
        L.xxm_share_candidates.argtypes = [vp, vp]
        L.xxm_maximize_motif_length_on.restype = i32
        L.xxm_maximize_motif_length_on.argtypes = [P(vp), i32, P(XxmMotif), P(XxmRefineParams), i32, P(i32)]
        L.xxm_iterate_motif_on.restype = i32
        L.xxm_iterate_motif_on.argtypes = [P(vp), i32, P(XxmMotif), P(XxmRefineParams), P(i32)]
        L._refine_ready = True
    return L


class Motif:
    """The part of class Motif (refinementPhase/Motif.h) the refinement loop reads and writes; sites are (sequence, position
    relative to PWM column 1) like Motif::_sites."""

    def __init__(self, pwm, cols, posSetSize, enr=(0, 0, 0), siteSeq=(), sitePos=(), siteCap=1 << 20, pval=0.0, posPval=0.0):
        self.m = XxmMotif()
        pwm = np.ascontiguousarray(pwm, dtype=np.float64).reshape(-1)
        assert pwm.size == 31 * 5
        C.memmove(self.m.pwm, pwm.ctypes.data, pwm.nbytes)
        cols = [int(c) for c in cols]
        self.m.ncols = len(cols)
        for i, c in enumerate(cols):
            self.m.cols[i] = c
        self.m.posSetSize = int(posSetSize)
        for i in range(3):
            self.m.enr[i] = int(enr[i])
        self.m.pval, self.m.posPval, self.m.worstScore = float(pval), float(posPval), 1.0
        self._seq = np.zeros(siteCap, dtype=np.int32)
        self._pos = np.zeros(siteCap, dtype=np.int32)
        n = len(siteSeq)
        self._seq[:n] = siteSeq
        self._pos[:n] = sitePos
        self.m.nsites, self.m.siteCap = n, siteCap
        self.m.siteSeq = self._seq.ctypes.data
        self.m.sitePos = self._pos.ctypes.data

    pwm = property(lambda s: np.array(s.m.pwm[:], dtype=np.float64))
    cols = property(lambda s: np.array(s.m.cols[:s.m.ncols], dtype=np.int32))
    enr = property(lambda s: np.array(s.m.enr[:], dtype=np.int32))
    pval = property(lambda s: s.m.pval)
    posPval = property(lambda s: s.m.posPval)
    worstScore = property(lambda s: s.m.worstScore)
    siteSeq = property(lambda s: s._seq[:s.m.nsites].copy())
    sitePos = property(lambda s: s._pos[:s.m.nsites].copy())


class Refiner:
    """IterateMotif + the Global:: switches it reads, bound to one context."""

    def __init__(self, ctx, totalPositions, revcomp, localization, maximizeMotifLength=True, minMatchPositions=5, maxMatchPositions=17,
                 pseudo=0.1, neff_pwm=10.0):
        self.ctx = ctx
        self.p = XxmRefineParams(int(totalPositions), int(revcomp), int(localization), int(maximizeMotifLength), int(minMatchPositions),
                                 int(maxMatchPositions), float(pseudo), float(neff_pwm), None)
        self._elong = np.zeros(31 * 5, dtype=np.float64)  # IterateMotif::elongPWM
        self.p.elongPWM = self._elong.ctypes.data

    def optimize_motif(self, motif):
        self.ctx._ck(_lib().xxm_optimize_motif(self.ctx.h, C.byref(motif.m), C.byref(self.p)))

    def maximize_motif_length(self, motif, lengthCorrection=None):
        nc = C.c_int32(0)
        lc = self.p.maximizeMotifLength if lengthCorrection is None else int(lengthCorrection)
        self.ctx._ck(_lib().xxm_maximize_motif_length(self.ctx.h, C.byref(motif.m), C.byref(self.p), lc, C.byref(nc)))
        return bool(nc.value)

    def iterate_motif(self, motif, views=()):
        """the per-motif loop of MotifContainer::iterate; views (refine.make_view of this context): the column variants of
        maximizeMotifLength are spread over this context and the views (xxm_iterate_motif_on), same result"""
        r = C.c_int32(0)
        if not views:
            self.ctx._ck(_lib().xxm_iterate_motif(self.ctx.h, C.byref(motif.m), C.byref(self.p), C.byref(r)))
            return r.value
        ctxs = [self.ctx] + list(views)
        hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        rc = _lib().xxm_iterate_motif_on(hs, len(ctxs), C.byref(motif.m), C.byref(self.p), C.byref(r))
        if rc:
            msgs = [c.L.xxm_last_error(c.h).decode() for c in ctxs]
            raise capi.XxmError("xxm_iterate_motif_on: rc %d: %s" % (rc, "; ".join(x for x in msgs if x) or "no message"))
        return r.value

    def maximize_motif_length_on(self, motif, views, lengthCorrection=None):
        nc = C.c_int32(0)
        lc = self.p.maximizeMotifLength if lengthCorrection is None else int(lengthCorrection)
        ctxs = [self.ctx] + list(views)
        hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        rc = _lib().xxm_maximize_motif_length_on(hs, len(ctxs), C.byref(motif.m), C.byref(self.p), lc, C.byref(nc))
        if rc:
            msgs = [c.L.xxm_last_error(c.h).decode() for c in ctxs]
            raise capi.XxmError("xxm_maximize_motif_length_on: rc %d: %s" % (rc, "; ".join(x for x in msgs if x) or "no message"))
        return bool(nc.value)


def make_view(owner):
    """A new context on the owner's GPU that reads the owner's resident set, options and NullModel tables and has its own stream,
    model and site lists (xxm_view_create).  Close the views before the owner."""
    L = _lib()
    h = C.c_void_p()
    if L.xxm_view_create(owner.h, C.byref(h)):
        raise capi.XxmError(L.xxm_last_error(None).decode())
    v = capi.Context.__new__(capi.Context)
    v.L, v.h, v.nent, v.ncols = L, h, owner.nent, 0
    return v


def share_seqset(ctx, owner):
    """xxm_share_seqset: an existing context becomes a view of owner's set."""
    ctx._ck(_lib().xxm_share_seqset(ctx.h, owner.h))
    ctx.nent = owner.nent


def view_count(ctx):
    return int(_lib().xxm_view_count(ctx.h))


def iterate_motifs(ctxs, motifs, params):
    """xxm_iterate_motifs: the motifs of MotifContainer::iterate several at a time, one host thread per context (an owner and its
    views).  motifs: list of Motif, updated in place; params: one Refiner (its switches are used for every motif).  Returns
    the rounds of every motif."""
    L = _lib()
    n = len(motifs)
    arr = (XxmMotif * max(1, n))()
    for i, m in enumerate(motifs):
        C.memmove(C.byref(arr[i]), C.byref(m.m), C.sizeof(XxmMotif))  # (the site arrays stay the Motif objects' own: only pointers move)
    hs = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rounds = (C.c_int32 * max(1, n))()
    bad, badCtx = C.c_int32(-1), C.c_int32(-1)
    p = XxmRefineParams()
    C.memmove(C.byref(p), C.byref(params.p), C.sizeof(XxmRefineParams))
    p.elongPWM = None
    rc = L.xxm_iterate_motifs(hs, len(ctxs), arr, n, C.byref(p), 1, rounds, C.byref(bad), C.byref(badCtx))
    for i, m in enumerate(motifs):
        C.memmove(C.byref(m.m), C.byref(arr[i]), C.sizeof(XxmMotif))
    if rc:
        who = ctxs[badCtx.value] if 0 <= badCtx.value < len(ctxs) else None
        msg = L.xxm_last_error(who.h).decode() if who is not None else "invalid arguments"
        raise capi.XxmError("motif %d: %s" % (bad.value, msg))
    return [int(rounds[i]) for i in range(n)]
