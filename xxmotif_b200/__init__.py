"""xxmotif_b200 -- B200 (sm_100a) implementation of XXmotif's PWM-refinement hot path.

The product is the C-ABI shared library libxxmotif_gpu.so (include/xxmotif_gpu.h, sources in xxmotif_b200/csrc);
`capi` is its ctypes face.  Importing the package never silently degrades: any use without the built library or
without a CUDA device raises.
"""
from .capi import (Context, XxmError, load_library, SITE_DTYPE, LIST_SORTED_SITES, LIST_SORTED_SITES_ELONG, LIST_BEST_STARTPOS,  # noqa: F401
                   ORDER_AS_IS, ORDER_BY_SCORE, ORDER_BY_STARTPOS, EXPORTS, LIB_PATH)
