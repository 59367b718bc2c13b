#!/usr/bin/env python
"""N-rank parity of the sharded refinement path on real GPUs, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/comm_parity.py [out.json]

For several workloads of the matrix (forward ZOOPS, --revcomp --mops, 12 match columns = two p-value tables, localization,
ragged lengths) every rank loads its block of the sequences into a context with an NCCL communicator attached and runs
  * three optimizeMotif steps (xxm_optimize_step: histogram exchange or run exchange, count all-reduce),
  * one candidate re-score under a longer column set + its subset choice and PWM recount (xxm_rescore,
    xxm_update_pwm_with_best_subset on the sharded lists),
  * the whole per-motif refinement loop (xxm_iterate_motif: optimizeMotif + maximizeMotifLength over all column variants
    until the p-value stops improving),
while rank 0 also runs the same calls on ONE context holding the whole set.  Everything the calls return must be identical:
global site counts, optimal subset size, log p-value and every PWM entry.  Rank 0 writes one JSON object; exit code 1 on a
mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import xxmotif_b200 as xb  # noqa: E402
from xxmotif_b200 import hostprep, workloads as W  # noqa: E402

CASES = [("cfg4", 24000), ("cfg2", 16000), ("cols12", 12000), ("cfg5", 30000), ("ragged", 16000), ("cfg3", 20000)]


def run_calls(ctx, w, n, codes, pwm0, cols, enr):
    """the calls under test on whatever ctx holds; returns a list of comparable results"""
    nent = len(n)
    total, _, _, _ = ctx.seqset_info()
    out = []
    pwm = pwm0.copy()
    for _ in range(3):
        pwm, info = ctx.optimize_step(pwm, cols, nent, int(total[nent]), enr, 0.1)
        out.append(("step", int(info["nbSites_global"]), int(info["bestMotifNb"]), float(info["bestPval"]), int(info["maxSites"]), pwm[5:].tobytes()))
    # one column variant of maximizeMotifLength: one more column at the end, candidate re-score + subset + recount
    tmp = np.append(cols, cols[-1] + 1).astype(np.int32)
    ctx.reinitialize(pwm, tmp, nent, int(total[nent]))
    ctx.rescore(int(cols[0]), nent, enr)
    upd, nb, bp, e1 = ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES_ELONG, pwm, int(tmp[0]), nent, True, 0.1)
    out.append(("rescore", int(upd), int(nb), float(bp), e1[5:].tobytes()))
    return out


def run_refinement(ctx, w, n, pwm0, cols, sharded):
    """the whole per-motif loop (xxm_iterate_motif: optimizeMotif + maximizeMotifLength with all column variants until the p-value
    stops improving) on whatever ctx holds; the site count is summed over the ranks"""
    from xxmotif_b200 import refine
    nent = len(n)
    total, _, mxl, _ = ctx.seqset_info()
    c0 = int(cols[0])
    raw_max = int(n.max())
    mot = refine.Motif(pwm0, cols, nent, enr=(1 - c0, 1 - c0, raw_max - c0 + 1), siteCap=1 << 20)
    R = refine.Refiner(ctx, totalPositions=int(total[nent]), revcomp=w.revcomp, localization=w.localization, maximizeMotifLength=True, minMatchPositions=5,
                       maxMatchPositions=17, pseudo=0.1, neff_pwm=10.0)
    rounds = R.iterate_motif(mot)
    ns = torch.tensor([int(mot.m.nsites)], device="cuda", dtype=torch.int64)
    if sharded:
        dist.all_reduce(ns)
    return ("refine", int(rounds), [int(c) for c in mot.cols], float(mot.pval), int(ns.item()), float(mot.m.worstScore)), float(mot.m.posPval)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = xb.Context(local)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.from_numpy(xb.capi.comm_unique_id()))
    dist.broadcast(idt, src=0)
    ctx.comm_init(rank, world, idt.cpu().numpy())
    report = {"ranks": world, "cases": [], "ok": True}
    for name, nseq in CASES:
        w = W.by_name(name).scaled(nseq, neg=8000 if W.by_name(name).neg else None)
        n, codes = w.positive_set()
        neg = w.negative_set(n, codes)
        bn, bc = neg if neg is not None else (n, codes)
        if w.revcomp:
            bn, bc = hostprep.revcomp_double(bn, bc)
        bg, bgl = hostprep.background(bn, bc)
        rp, rc = hostprep.null_model(bn, bc, order=w.order)
        st = w.start_pwm()
        if st is None:  # -m start: the IUPAC PWM of the TATA profile stands in (any PWM is a valid input of the path)
            st = hostprep.profile_pwm()
            if "." in w.start[1]:
                pw, cl = st
                st = (pw, np.array([c for c in cl if c != cl[5]], dtype=np.int32))  # a gap column like TATAW.AWR
        pwm0, cols = st
        c0 = int(cols[0])
        enr = (1 - c0 + 1, int(n.max()) - c0 + 1)
        off = np.concatenate(([0], np.cumsum(n.astype(np.int64))))
        # contiguous blocks of sequences, deliberately uneven
        cuts = [0] + [int(nseq * (r + 1) / world * (0.9 if r % 2 == 0 and r + 1 < world else 1.0)) for r in range(world)]
        cuts[-1] = nseq
        lo, hi = cuts[rank], cuts[rank + 1]
        ctx.set_options(mops=w.mops, oops=False, localization=w.localization, bg=bg, bgLog=bgl)
        ctx.set_background(w.order, rp, rc)
        ctx.load_seqset(n[lo:hi], codes[off[lo]:off[hi]], revcomp=w.revcomp)
        err = None
        pp = None
        try:
            got = run_calls(ctx, w, n[lo:hi], codes[off[lo]:off[hi]], pwm0, cols, enr)
            r, pp = run_refinement(ctx, w, n[lo:hi], pwm0, cols, True)
            got.append(r)
        except xb.XxmError as ex:
            got, err = None, str(ex)
        entry = {"workload": w.describe(), "sequences": nseq, "match_columns": int(len(cols))}
        bad = torch.zeros(1, device="cuda", dtype=torch.int32)
        if rank == 0:
            solo = xb.Context(local)
            try:
                solo.set_options(mops=w.mops, oops=False, localization=w.localization, bg=bg, bgLog=bgl)
                solo.set_background(w.order, rp, rc)
                solo.load_seqset(n, codes, revcomp=w.revcomp)
                want = run_calls(solo, w, n, codes, pwm0, cols, enr)
                r1, pp1 = run_refinement(solo, w, n, pwm0, cols, False)
                want.append(r1)
            finally:
                solo.close()
            same = got is not None and got == want
            if pp is not None and pp1 is not None:  # the positional log sum is added up in a different order: equal to rounding
                same = same and abs(pp - pp1) <= 1e-9 * max(1.0, abs(pp1))
                entry["posPval"] = [pp, pp1]
            entry.update({"identical": bool(same), "error": err, "single_gpu": [list(x[1:5]) if x[0] == "refine" else list(x[1:4]) for x in want],
                          "sharded": [list(x[1:5]) if x[0] == "refine" else list(x[1:4]) for x in got] if got else None})
            bad[0] = 0 if same else 1
        dist.broadcast(bad, src=0)
        if rank == 0:
            report["cases"].append(entry)
            report["ok"] = report["ok"] and entry["identical"]
    ctx.close()
    if rank == 0:
        txt = json.dumps(report, indent=1)
        print(txt)
        if len(sys.argv) > 1:
            with open(sys.argv[1], "w") as fh:
                fh.write(txt + "\n")
    dist.destroy_process_group()
    sys.exit(0 if report["ok"] or rank != 0 else 1)


if __name__ == "__main__":
    main()
