#!/usr/bin/env python
"""SURVEY 8(f) row 2, variant half: the whole refinement of ONE motif (xxm_iterate_motif: optimizeMotif + maximizeMotifLength until
the p-value stops improving) with the <= 11 column variants of maximizeMotifLength evaluated one after the other on one context
(the reference's loop) and spread over the context and its views (xxm_iterate_motif_on).  Same workload as bench.py's refine_run
(-p TATA profile, ZOOPS, N x 1 kb).  One JSON line per size; results must be identical.

    python tools/variant_batch.py [--sizes 100000,1000000] [--contexts 1,2,3,4]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_size(nseq, seqlen, nctxs, repeats):
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, refine
    n, codes = hostprep.synthetic_set(nseq, seqlen, seed=4)
    sample = min(nseq, 100000)  # background and NullModel from a prefix: inputs of the path, any consistent model does
    bg, bgl = hostprep.background(n[:sample], codes[:sample * seqlen])
    rp, rc = hostprep.null_model(n[:sample], codes[:sample * seqlen], order=2)
    pwm0, cols = hostprep.profile_pwm()
    owner = xb.Context(0)
    views = []
    try:
        owner.set_options(mops=False, bg=bg, bgLog=bgl)
        owner.load_seqset(n, codes, revcomp=False)
        owner.set_background(2, rp, rc)
        total, _, mxl, _ = owner.seqset_info()
        c0 = int(cols[0])
        R = refine.Refiner(owner, totalPositions=int(total[nseq]), revcomp=False, localization=False)

        def motif():
            return refine.Motif(pwm0, cols, nseq, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=max(1 << 20, 2 * nseq))

        def state(m):
            return (m.cols.tolist(), m.pwm.tobytes(), m.pval, m.posPval, m.siteSeq.tobytes(), m.sitePos.tobytes())

        views = [refine.make_view(owner) for _ in range(max(nctxs) - 1)]
        res, ref = {}, None
        for k in nctxs:
            secs = []
            for _ in range(repeats + 1):  # the first run on a set of contexts also pays for the buffers they allocate
                m = motif()
                t0 = time.perf_counter()
                rounds = R.iterate_motif(m, views=views[:k - 1])
                secs.append(time.perf_counter() - t0)
            st = state(m)
            ref = st if ref is None else ref
            res[str(k)] = {"seconds": min(secs[1:]), "first_run_seconds": secs[0], "identical_to_one_context": st == ref, "rounds": int(rounds)}
        return {"sequences": nseq, "length": seqlen, "sites": int(m.m.nsites), "log_pval": float(m.pval), "match_columns": [int(m.cols[0]), int(m.cols[-1])],
                "contexts": res, "ok": all(v["identical_to_one_context"] for v in res.values())}
    finally:
        for v in views:
            v.close()
        owner.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="100000,1000000")
    ap.add_argument("--seqlen", type=int, default=1000)
    ap.add_argument("--contexts", default="1,2,3,4")
    ap.add_argument("--repeats", type=int, default=2)
    a = ap.parse_args()
    ok = True
    for s in a.sizes.split(","):
        r = run_size(int(s), a.seqlen, [int(x) for x in a.contexts.split(",")], a.repeats)
        r["what"] = "xxm_iterate_motif of the TATA profile, column variants on 1..k contexts of one GPU"
        ok = ok and r["ok"]
        print(json.dumps(r), flush=True)
    return 0 if ok else 3


if __name__ == "__main__":
    sys.exit(main())
