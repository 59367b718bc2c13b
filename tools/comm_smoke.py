#!/usr/bin/env python
"""Two-rank smoke test of the native sharded step (xxm_comm_init + xxm_optimize_step over NCCL) against the split-phase
path with the exchange in Python.  torchrun --nproc-per-node 2 tools/comm_smoke.py [--nseq N]"""
import argparse
import faulthandler
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(60, exit=True)
os.environ.setdefault("NCCL_DEBUG", "WARN")
import numpy as np
import torch
import torch.distributed as dist

import xxmotif_b200 as xb
from xxmotif_b200 import hostprep, sharded

ap = argparse.ArgumentParser()
ap.add_argument("--nseq", type=int, default=20000)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])


def log(*x):
    print("[rank %d]" % rank, *x, file=sys.stderr, flush=True)


torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, codes = hostprep.synthetic_set(a.nseq, 500, seed=4 + rank)
bg, bgl = hostprep.background(n[:2000], codes[:2000 * 500])
rp, rc = hostprep.null_model(n[:2000], codes[:2000 * 500], order=2)
for arr in (bg, bgl, rp, rc):  # one background model for the whole set
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    dist.broadcast(t, src=0)
    arr[...] = t.cpu().numpy()
pwm0, cols = hostprep.profile_pwm()
res = {}
for native in (False, True):
    ctx = xb.Context(local)
    ctx.set_options(mops=False, bg=bg, bgLog=bgl)
    ctx.set_background(2, rp, rc)
    ctx.load_seqset(n, codes, revcomp=False)
    total, avg, mxl, mnl = ctx.seqset_info()
    log("native", native, "creating job")
    job = sharded.ShardedRefinement(ctx, dist, nent_local=len(n), total_local=int(total[len(n)]), pseudo=0.1, bg=bg, native=native)
    log("job ready")
    enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
    pwm = pwm0.copy()
    for it in range(3):
        pwm, info = job.optimize_step(pwm, cols, enr)
        log("native=%s step %d K=%d bestPval=%r" % (native, it, info["bestMotifNb"], info["bestPval"]))
    res[native] = (pwm, info["bestMotifNb"], info["bestPval"])
    torch.cuda.synchronize()
    ctx.close()
    log("closed")
ok = (res[False][0] == res[True][0]).all() and res[False][1:] == res[True][1:]
log("native == split-phase:", bool(ok))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
