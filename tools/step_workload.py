#!/usr/bin/env python
"""Run a few optimizeMotif steps of one workload of the matrix on cuda:0 (for ncu launch lists / captures of one configuration):

    python tools/step_workload.py cfg2 1000000 3
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402


def main():
    name, nseq, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    import xxmotif_b200 as xb
    from xxmotif_b200 import workloads as W
    w = W.by_name(name)
    ref = bench.run_reference_workload(w.scaled(1500, neg=5000 if w.neg else None), 1, 0)
    wl = w.scaled(nseq)
    n, codes = wl.positive_set()
    neg = wl.negative_set(n, codes)
    ctx = xb.Context(0)
    t0 = time.perf_counter()
    r = bench.time_workload(ctx, wl, n, codes, neg, ref, steps, 6556.5)
    print(name, nseq, "ms/step %.3f" % r["ms_per_step"], "value %.3e" % r["value"], r["scan"], r["host_phases_us"], "total %.1f s" % (time.perf_counter() - t0))
    if os.environ.get("XXM_TRACE"):
        from xxmotif_b200 import capi
        capi.load_library().xxm_trace_dump()
    ctx.close()


if __name__ == "__main__":
    main()
