#!/usr/bin/env python
"""SURVEY 8(f) row 2, "batch motifs": the motifs of MotifContainer::iterate refined one after the other on one context (the
reference's loop) and several at a time on views of the same resident set (xxm_iterate_motifs), on the small sets of
BASELINE's configs 1, 2 and 5 where one motif leaves most of a B200 idle.  Prints one JSON line; results must be identical.

    python tools/motif_batch.py [--contexts 4] [--motifs 16] [--shapes cfg1,cfg2]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {  # name: (sequences, length, revcomp, mops)
    "cfg1": (500, 200, False, False),
    "cfg2": (10000, 500, True, True),
    "cfg5": (50000, 300, False, False),
}


def start_motifs(nmotifs, nent, mxl):
    """distinct valid start models: TATA profiles of 5..12 columns and GC-rich ones, repeated with the columns in other places
    (a discovery run hands the refinement a few hundred seeds of this kind)"""
    from xxmotif_b200 import hostprep, refine, workloads
    profs = [workloads.profile(k) for k in (8, 6, 7, 10, 12, 5, 9, 11)]
    gc = np.array([[5, 5, 10, 5, 5, 10], [45, 40, 70, 45, 5, 40], [45, 50, 10, 45, 85, 40], [5, 5, 10, 5, 5, 10]], dtype=np.float64)
    profs += [gc, gc[:, ::-1].copy(), np.roll(gc, 1, axis=0), np.roll(gc, 2, axis=0)]
    rng = np.random.default_rng(5)
    while len(profs) < nmotifs:  # sharpened / flattened copies
        p = profs[int(rng.integers(0, 12))].copy()
        p = p ** float(rng.uniform(0.6, 1.4))
        profs.append(100.0 * p / p.sum(axis=0, keepdims=True))
    out = []
    for prof in profs[:nmotifs]:
        pwm, cols = hostprep.profile_pwm(prof)
        c0 = int(cols[0])
        out.append(lambda pwm=pwm, cols=cols, c0=c0: refine.Motif(pwm, cols, nent, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=max(1 << 20, 2 * nent)))
    return out


def state(m):
    return (m.cols.tolist(), m.pwm.tobytes(), m.enr.tolist(), m.pval, m.posPval, m.siteSeq.tobytes(), m.sitePos.tobytes())


def run_shape(name, nctx, nmotifs, repeats=2):
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, refine
    nseq, length, revcomp, mops = SHAPES[name]
    n, codes = hostprep.synthetic_set(nseq, length, seed=21)
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    owner = xb.Context(0)
    views = []
    try:
        owner.set_options(mops=mops, bg=bg, bgLog=bgl)
        owner.load_seqset(n, codes, revcomp=revcomp)
        owner.set_background(2, rp, rc)
        total, _, mxl, _ = owner.seqset_info()
        R = refine.Refiner(owner, totalPositions=int(total[nseq]), revcomp=revcomp, localization=False)
        make = start_motifs(nmotifs, nseq, mxl)
        views = [refine.make_view(owner) for _ in range(nctx - 1)]
        ctxs = [owner] + views
        # warm every context (buffers are allocated on first use; a real run refines hundreds of motifs per context)
        refine.iterate_motifs(ctxs, [make[i % len(make)]() for i in range(len(ctxs))], R)
        serial_s, batch_s, ref_state = [], [], None
        for _ in range(repeats):
            ms = [mk() for mk in make]
            t0 = time.perf_counter()
            for m in ms:
                R.iterate_motif(m)
            serial_s.append(time.perf_counter() - t0)
            st = [state(m) for m in ms]
            ms = [mk() for mk in make]
            t0 = time.perf_counter()
            refine.iterate_motifs(ctxs, ms, R)
            batch_s.append(time.perf_counter() - t0)
            same = [state(m) for m in ms] == st
            ref_state = same if ref_state is None else (ref_state and same)
        return {"shape": name, "sequences": nseq, "length": length, "revcomp": revcomp, "mops": mops, "motifs": nmotifs, "contexts": nctx,
                "serial_seconds": min(serial_s), "batched_seconds": min(batch_s), "speedup": min(serial_s) / min(batch_s), "identical": bool(ref_state)}
    finally:
        for v in views:
            v.close()
        owner.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contexts", type=int, default=4)
    ap.add_argument("--motifs", type=int, default=16)
    ap.add_argument("--shapes", default="cfg1,cfg2")
    a = ap.parse_args()
    res = []
    for name in a.shapes.split(","):
        res.append(run_shape(name, a.contexts, a.motifs))
    out = {"what": "motifs of MotifContainer::iterate: one context, one after the other (the reference's loop) vs xxm_iterate_motifs on views of the same set",
           "host_cores": os.cpu_count(), "results": res, "ok": all(r["identical"] for r in res)}
    print(json.dumps(out))
    return 0 if out["ok"] else 3


if __name__ == "__main__":
    sys.exit(main())
