#!/usr/bin/env python
"""Wall time of xxm_load_seqset (host packing + chunked copy) for several packing-thread counts.
   python tools/load_probe.py [--nseq N] [--seqlen L] [--revcomp]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import xxmotif_b200 as xb
from xxmotif_b200 import hostprep

ap = argparse.ArgumentParser()
ap.add_argument("--nseq", type=int, default=1000000)
ap.add_argument("--seqlen", type=int, default=1000)
ap.add_argument("--revcomp", action="store_true")
a = ap.parse_args()
n, codes = hostprep.synthetic_set(a.nseq, a.seqlen, seed=4)
ctx = xb.Context(0)
ctx.load_seqset(n, codes, revcomp=a.revcomp)
for t in (1, 2, 4, 8, 12, 16, 24, 32):
    if t > 2 * (os.cpu_count() or 1):
        break
    os.environ["XXM_PACK_THREADS"] = str(t)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.load_seqset(n, codes, revcomp=a.revcomp)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("threads %2d: %.2f ms  (%.1f GB/s of raw bases)" % (t, best * 1e3, codes.size / best / 1e9), flush=True)
# packing alone, one thread, no copies
t0 = time.perf_counter()
hostprep.pack_seqset(n[:100000], codes[:int(n[:100000].sum())], a.revcomp)
t0 = time.perf_counter()
hostprep.pack_seqset(n[:100000], codes[:int(n[:100000].sum())], a.revcomp)
dt = time.perf_counter() - t0
print("pack only, 1 thread, 100k sequences: %.2f ms (%.1f GB/s)" % (dt * 1e3, int(n[:100000].sum()) / dt / 1e9))
