#!/usr/bin/env python
"""Wall time of xxm_load_seqset on 1M x 1 kb through each packing path (XXM_LOAD_PATH) for a few thread counts, inputs in pinned
host memory as in bench.py's e2e leg:   XXM_TRACE=1 python tools/load_paths.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import xxmotif_b200 as xb
from xxmotif_b200 import hostprep

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n, codes = hostprep.synthetic_set(nseq, 1000, seed=4)
pinned = torch.from_numpy(codes).pin_memory()
codes = pinned.numpy()
ctx = xb.Context(0)
out = {"what": "xxm_load_seqset wall ms, %d x 1000 bp, pinned input, cpu_count %d" % (nseq, os.cpu_count() or 0), "runs": []}
cases = [("host", 12), ("host", 6), ("device", 1), ("hybrid", 1), ("hybrid", 2), ("hybrid", 3), ("hybrid", 4), ("hybrid", 5), ("host", 3)]
if len(sys.argv) > 2:
    cases = [tuple(c.split(":")) for c in sys.argv[2:]]
for case in cases:
    path, threads = case[0], int(case[1])
    os.environ["XXM_LOAD_PATH"] = path
    os.environ["XXM_PACK_THREADS"] = str(threads)
    if len(case) > 2:  # path:threads:chunkKB
        os.environ["XXM_PACK_CHUNK_KB"] = case[2]
    else:
        os.environ.pop("XXM_PACK_CHUNK_KB", None)
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.load_seqset(n, codes)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    r = {"path": path, "threads": threads, "chunk_kb": int(case[2]) if len(case) > 2 else None, "ms": round(best * 1e3, 2), "GBs_raw": round(codes.size / best / 1e9, 1)}
    print(r, flush=True)
    out["runs"].append(r)
print(json.dumps(out))
