#!/usr/bin/env python
"""Per-phase wall times of one refinement step through the C ABI (used to decide what to optimise; also the command
that is profiled under ncu).   python tools/phase_times.py [--nseq N] [--seqlen L] [--steps K]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import xxmotif_b200 as xb
from xxmotif_b200 import hostprep, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--nseq", type=int, default=100000)
ap.add_argument("--seqlen", type=int, default=1000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--mops", action="store_true")
ap.add_argument("--revcomp", action="store_true")
a = ap.parse_args()

n, codes = hostprep.synthetic_set(a.nseq, a.seqlen, seed=4)
if a.revcomp:
    n2, c2 = hostprep.revcomp_double(n, codes)
else:
    n2, c2 = n, codes
bg, bgl = hostprep.background(n2, c2)
rp, rc = hostprep.null_model(n2, c2, order=2)
ctx = xb.Context(0)
ctx.set_options(mops=a.mops, bg=bg, bgLog=bgl)
ctx.set_background(2, rp, rc)
t0 = time.perf_counter()
ctx.load_seqset(n, codes, revcomp=a.revcomp)
print("load_seqset %.1f ms" % ((time.perf_counter() - t0) * 1e3))
total, avg, mxl, mnl = ctx.seqset_info()
pwm, cols = hostprep.profile_pwm()
nent = len(n)
enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
span = int(cols[-1] - cols[0] + 1)


def sync():
    torch.cuda.synchronize()


for it in range(3 + a.steps):
    T = {}
    sync(); t = time.perf_counter()
    ctx.reinitialize(pwm, cols, nent, int(total[nent])); sync()
    T["reinit"] = time.perf_counter() - t; t = time.perf_counter()
    nb, nbest, mx = ctx.scan_full(nent, enr); sync()
    T["scan_full"] = time.perf_counter() - t; t = time.perf_counter()
    sc, cn = sharded.score_runs(ctx, xb.LIST_SORTED_SITES); sync()
    T["runs"] = time.perf_counter() - t; t = time.perf_counter()
    K, bp, thr = sharded.choose_subset(sc, cn, nb, nent * 1000, int(total[nent]) - (2 if a.revcomp else 1) * nent * (span - 1))
    T["choose"] = time.perf_counter() - t; t = time.perf_counter()
    counts = sharded.count_sites(ctx, xb.LIST_SORTED_SITES, thr, int(cols[0])); sync()
    T["count"] = time.perf_counter() - t; t = time.perf_counter()
    pwm = sharded.build_pwm(counts, 0.1 * K, bg)
    T["pwm"] = time.perf_counter() - t
    st = ctx.stats()
    if it >= 3:
        print("step %d: " % it + " ".join("%s=%.3fms" % (k, v * 1e3) for k, v in T.items()) +
              " | k_scan=%.3fms hits=%d kept=%d K=%d runs=%d" % (st["last_scan_ms"], nbest, nb, K, len(sc)))
ctx.close()
