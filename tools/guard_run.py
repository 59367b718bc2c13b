#!/usr/bin/env python
"""Run the golden replays, the ragged / low-complexity stress case and small matrix workloads with a canary behind every device
buffer (XXM_DEBUG_GUARDS=1, set before the library is loaded) and report overwritten canaries as JSON.  compute-sanitizer is
closed on this GPU pool; this is the in-house substitute for its memcheck pass over writes past the end of a buffer."""
import json
import os
import sys

os.environ["XXM_DEBUG_GUARDS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import xxmotif_b200 as xb  # noqa: E402
from xxmotif_b200 import hostprep, workloads as W  # noqa: E402
from tracefile import read_trace  # noqa: E402
from replay import TraceInputs, replay  # noqa: E402
from gpu_backend import GpuBackend  # noqa: E402

report = {"cases": [], "checked": 0, "damaged": 0}


def check(ctx, name):
    c, d = ctx.check_guards()
    report["cases"].append({"case": name, "buffers": c, "damaged": d})
    report["checked"] += c
    report["damaged"] += d


gold = os.path.join(ROOT, "tests", "golden")
for f in sorted(os.listdir(gold)):
    if not f.endswith(".trace.gz"):
        continue
    t = read_trace(os.path.join(gold, f))
    be = GpuBackend(TraceInputs(t))
    replay(t, be)
    check(be.ctx, f[:-len(".trace.gz")])
    be.ctx.close()

# ragged, low-complexity sequences (dense hit clusters, tiles that overflow the expand queue), both strands, a few steps
import test_gpu_scale as tgs  # noqa: E402
n, codes = tgs._low_complexity_set()
for revcomp, mops in ((False, False), (True, True)):
    n2, c2 = hostprep.revcomp_double(n, codes) if revcomp else (n, codes)
    bg, bgl = hostprep.background(n2, c2)
    rp, rc = hostprep.null_model(n2, c2, order=2)
    ctx = xb.Context(0)
    ctx.set_options(mops=mops, bg=bg, bgLog=bgl)
    ctx.set_background(2, rp, rc)
    ctx.load_seqset(n, codes, revcomp=revcomp)
    total, _, mxl, _ = ctx.seqset_info()
    pwm, cols = hostprep.profile_pwm()
    enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
    for _ in range(3):
        pwm, info = ctx.optimize_step(pwm, cols, len(n), int(total[len(n)]), enr, 0.1)
    ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_BY_SCORE, cap=info["nbBestSites"] + 8)
    check(ctx, "low_complexity revcomp=%d mops=%d" % (revcomp, mops))
    ctx.close()

# the matrix workloads at a few thousand sequences: steps + a re-score
for w, _ in W.MATRIX:
    ws = w.scaled(6000 if w.name != "cols17" else 2500, neg=5000 if w.neg else None)
    n, codes = ws.positive_set()
    neg = ws.negative_set(n, codes)
    bn, bc = neg if neg is not None else (n, codes)
    if ws.revcomp:
        bn, bc = hostprep.revcomp_double(bn, bc)
    bg, bgl = hostprep.background(bn, bc)
    rp, rc = hostprep.null_model(bn, bc, order=ws.order)
    st = ws.start_pwm() or hostprep.profile_pwm()
    pwm, cols = st
    ctx = xb.Context(0)
    ctx.set_options(mops=ws.mops, localization=ws.localization, bg=bg, bgLog=bgl)
    ctx.set_background(ws.order, rp, rc)
    ctx.load_seqset(n, codes, revcomp=ws.revcomp)
    total, _, mxl, _ = ctx.seqset_info()
    c0 = int(cols[0])
    enr = (1 - c0 + 1, int(n.max()) - c0 + 1)
    for _ in range(2):
        pwm, info = ctx.optimize_step(pwm, cols, len(n), int(total[len(n)]), enr, 0.1)
    tmp = np.append(cols, cols[-1] + 1).astype(np.int32)
    ctx.reinitialize(pwm, tmp, len(n), int(total[len(n)]))
    ctx.rescore(c0, len(n), enr)
    ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES_ELONG, pwm, int(tmp[0]), len(n), True, 0.1)
    check(ctx, "matrix " + ws.name)
    ctx.close()

print(json.dumps(report))
sys.exit(0 if report["damaged"] == 0 else 1)
