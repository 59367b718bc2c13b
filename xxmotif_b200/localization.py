"""ctypes face of the localization statistics: enriched start-position region and the
positional p-value of a site list."""
import ctypes as C

import numpy as np

from . import capi

EXTRA_EXPORTS = ["xxm_update_enriched_region", "xxm_host_positional_pval", "xxm_positional_pval"]


def _lib():
    L = capi.load_library()
    if not getattr(L, "_loc_ready", False):
        vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
        L.xxm_update_enriched_region.restype = i32
        L.xxm_update_enriched_region.argtypes = [vp, i32, i32, vp]
        L.xxm_host_positional_pval.restype = i32
        L.xxm_host_positional_pval.argtypes = [i32, i32, i32, i32, i32, vp, i32, C.POINTER(dbl)]
        L._loc_ready = True
    return L


def update_enriched_region(ctx, firstMotifColumn, motifLength, enr):
    """enr = (max, startRegion, endRegion) -> new (max, startRegion, endRegion)"""
    e = np.ascontiguousarray(enr, dtype=np.int32).copy()
    ctx._ck(_lib().xxm_update_enriched_region(ctx.h, firstMotifColumn, motifLength, e.ctypes.data_as(C.c_void_p)))
    return int(e[0]), int(e[1]), int(e[2])


def positional_pval(max_leng, revcomp, enrMax, firstMotifColumn, motifLength, sitePos):
    sp = np.ascontiguousarray(sitePos, dtype=np.int32)
    out = C.c_double(0.0)
    rc = _lib().xxm_host_positional_pval(max_leng, int(revcomp), enrMax, firstMotifColumn, motifLength, sp.ctypes.data_as(C.c_void_p), len(sp), C.byref(out))
    if rc:
        raise capi.XxmError("positional p-value undefined (the reference exits: in no man's land)")
    return out.value
