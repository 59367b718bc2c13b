#!/usr/bin/env python
"""ONE host thread, N GPUs (xxm_multi): the calls of the reference-side binding over the set of the metric, timed by wall clock
around the call a caller makes (reinitialize + full scan + both lists back on the host), for N = 1 and every N given.

    python tools/multi_scan.py 1000000 1000 1 2 4 8      # sequences, length, device counts
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    nseq, length = int(sys.argv[1]), int(sys.argv[2])
    devices = [int(a) for a in sys.argv[3:]] or [1]
    from xxmotif_b200 import capi, hostprep
    n, codes = hostprep.synthetic_set(nseq, length)
    bg, bgl = hostprep.background(n[:20000], codes[:int(n[:20000].sum())])
    rp, rc = hostprep.null_model(n[:20000], codes[:int(n[:20000].sum())], order=2)
    pwm, cols = hostprep.profile_pwm()
    total = int(n.sum())
    enr = (1 - int(cols[0]) + 1, length - int(cols[0]) + 1)
    out = {"what": "xxm_multi: one host thread, N devices; wall clock per call on %d x %d" % (nseq, length), "runs": []}
    base = None
    for nd in devices:
        ctx = capi.MultiContext(nd)
        ctx.set_options(bg=bg, bgLog=bgl)
        ctx.set_background(2, rp, rc)
        t0 = time.perf_counter()
        ctx.load_seqset(n, codes)
        load_s = time.perf_counter() - t0
        reps = 6
        ts, tl = [], []
        for i in range(reps + 2):
            t0 = time.perf_counter()
            ctx.reinitialize(pwm, cols, nseq, total)
            counts = ctx.scan_full(nseq, enr)
            t1 = time.perf_counter()
            best = ctx.get_sites(capi.LIST_BEST_STARTPOS, capi.ORDER_BY_SCORE, cap=counts[1] + 1)
            srt = ctx.get_sites(capi.LIST_SORTED_SITES, cap=counts[0] + 1)
            t2 = time.perf_counter()
            if i >= 2:
                ts.append(t1 - t0)
                tl.append(t2 - t1)
        sig = (counts, int(best["sequence"].astype(np.int64).sum()), int(srt["startPos"].astype(np.int64).sum()), float(best["score"].astype(np.float64).sum()))
        if base is None:
            base = sig
        run = {"devices": nd, "load_s": round(load_s, 3), "reinitialize_scan_ms": round(1000 * min(ts), 3), "lists_to_host_ms": round(1000 * min(tl), 3),
               "nbSites": counts[0], "nbBestSites": counts[1], "maxSites": counts[2], "same_as_first_run": sig == base,
               "windows_per_s_scan": (total - 7 * nseq) / min(ts)}
        print(run, flush=True)
        out["runs"].append(run)
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
