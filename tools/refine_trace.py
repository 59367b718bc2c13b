#!/usr/bin/env python
"""The whole refinement of the bench motif (xxm_iterate_motif on 1M x 1 kb) once for warm-up and once with XXM_TRACE call times:
where do the ~0.1 s go?   XXM_TRACE=1 python tools/refine_trace.py [nseq]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import xxmotif_b200 as xb  # noqa: E402
from xxmotif_b200 import hostprep, refine  # noqa: E402

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n, codes = hostprep.synthetic_set(nseq, 1000, seed=4)
bg, bgl = hostprep.background(n[:50000], codes[:50000000])
rp, rc = hostprep.null_model(n[:50000], codes[:50000000], order=2)
ctx = xb.Context(0)
ctx.set_options(mops=False, bg=bg, bgLog=bgl)
ctx.set_background(2, rp, rc)
ctx.load_seqset(n, codes)
pw, cl = hostprep.profile_pwm()
tt, _, mm, _ = ctx.seqset_info()
for _ in range(5):  # a few refinement steps first, as in bench.py: they pay for the LOG_sum_i checkpoints up to 10^9
    pw, _ = ctx.optimize_step(pw, cl, nseq, int(tt[nseq]), (1 - int(cl[0]) + 1, mm - int(cl[0]) + 1), 0.1)
for rep in range(3):
    print("---- run %d" % rep, file=sys.stderr)
    xb.capi.load_library().xxm_trace_dump()
    total, _, mxl, _ = ctx.seqset_info()
    pwm0, cols = hostprep.profile_pwm()
    c0 = int(cols[0])
    mot = refine.Motif(pwm0, cols, nseq, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=max(1 << 20, 2 * nseq))
    R = refine.Refiner(ctx, totalPositions=int(total[nseq]), revcomp=False, localization=False, maximizeMotifLength=True, minMatchPositions=5, maxMatchPositions=17,
                       pseudo=0.1, neff_pwm=10.0)
    torch.cuda.synchronize()
    l0 = ctx.stats()["launches"]
    t0 = time.perf_counter()
    rounds = R.iterate_motif(mot)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("run %d: %.4f s, %d rounds, %d launches, cols %d..%d, %d sites, log p %.6f" % (rep, dt, rounds, ctx.stats()["launches"] - l0, mot.cols[0], mot.cols[-1],
                                                                                      mot.m.nsites, mot.pval), file=sys.stderr)
ctx.close()  # prints the call times of the last run
