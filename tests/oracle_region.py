"""ctypes binding of the a15 restatements in oracle/xxm_oracle_region.c (test infrastructure)."""
import ctypes as C

import numpy as np

import oracle_lib as ol


def _lib():
    L = ol.lib()
    if not getattr(L, "_region_ready", False):
        vp, i32 = C.c_void_p, C.c_int32
        L.xo_update_enriched_region.restype = None
        L.xo_update_enriched_region.argtypes = [vp, vp, i32, i32, i32, i32, vp]
        L._region_ready = True
    return L


def update_enriched_region(octx, best_sorted, revcomp, max_leng, firstMotifColumn, motifLength, enr):
    """best_sorted: structured site array sorted by score (a terminating score-1 entry is appended here)."""
    b = np.zeros(len(best_sorted) + 1, dtype=ol.SITE_DTYPE)
    b[:len(best_sorted)] = best_sorted
    b[-1]["score"] = 1.0
    e = np.ascontiguousarray(enr, dtype=np.int32).copy()
    _lib().xo_update_enriched_region(octx.h, b.ctypes.data_as(C.c_void_p), int(revcomp), max_leng, firstMotifColumn, motifLength, e.ctypes.data_as(C.c_void_p))
    return int(e[0]), int(e[1]), int(e[2])
