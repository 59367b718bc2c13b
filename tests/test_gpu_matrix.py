"""GPU parity over the workload matrix: the CUDA path through the C ABI (xxm_optimize_step, xxm_update_enriched_region) against
the UNMODIFIED reference run on the same >= 10 000 sequences here (oracle/_ref/ref_harness bench), for the shapes of
BASELINE.json's configs 2-5, for 12 and 17 match columns and for a ragged set.  Four optimizeMotif steps must end on the
reference's site counts, optimal subset, log p-value, PWM and enriched region, bit for bit."""
import numpy as np
import pytest

import matrix_common as mc
from xxmotif_b200 import workloads as W

pytestmark = pytest.mark.gpu

STEPS = 4
SIZE = {"cfg2": 10000, "cfg3": 20000, "cfg4": 10000, "cfg5": 50000, "cols12": 10000, "cols17": 10000, "ragged": 10000}


def gpu_steps(w, ref, n, codes, neg, steps, ctx=None):
    """`steps` optimizeMotif iterations on the GPU from the reference's start motif; returns the harness-shaped dict"""
    import xxmotif_b200 as xb
    from xxmotif_b200 import localization
    bg, bgl, rp, rc = mc.background_inputs(w, n, codes, neg)
    own = ctx is None
    if own:
        ctx = xb.Context(0)
    try:
        ctx.set_options(mops=w.mops, oops=False, localization=w.localization, bg=bg, bgLog=bgl)
        ctx.set_background(w.order, rp, rc)
        ctx.load_seqset(n, codes, revcomp=w.revcomp)
        total, _, mxl, _ = ctx.seqset_info()
        nent = len(n)
        assert (nent, int(total[nent])) == (ref["posSetSize"], ref["total"])
        cols = np.asarray(ref["cols"], dtype=np.int32)
        span = int(cols[-1] - cols[0] + 1)
        pwm = np.asarray(ref["pwm0"], dtype=np.float64)
        enr = tuple(ref["enr0"])
        got = {}
        for _ in range(steps):
            pwm, info = ctx.optimize_step(pwm, cols, nent, int(total[nent]), (enr[1], enr[2]), 0.1)
            got.update(nbSites=info["nbSites"], nbBestSites=info["nbBestSites"], maxSites=info["maxSites"], bestMotifNb=info["bestMotifNb"],
                       bestPval=info["bestPval"])
            if w.localization:
                enr = localization.update_enriched_region(ctx, int(cols[0]), span, enr)
        got["pwm"] = pwm
        got["enr"] = enr
        return got
    finally:
        if own:
            ctx.close()


@pytest.mark.parametrize("name", [w.name for w, _ in W.MATRIX])
def test_gpu_steps_match_reference(name):
    if not mc.have_reference():
        pytest.skip("oracle/_ref/ref_harness not built")
    w = W.by_name(name).scaled(SIZE[name], neg=30000 if W.by_name(name).neg else None)
    ref, n, codes, neg = mc.run_reference(w, STEPS)
    got = gpu_steps(w, ref, n, codes, neg, STEPS)
    mc.compare_with_reference(ref, got)


@pytest.mark.parametrize("name,nseq", [("cfg5", 120000), ("cols12", 80000), ("cols17", 30000)])
def test_sweep_prefilter_equals_full_sweep(name, nseq):
    """Lists whose scores are not table valued (localization; two p-value tables) choose their subset from 10^4..10^6 runs of
    equal scores.  The device prefilter (k_sweep_eval: approximate evaluation of every run, exact evaluation of the ones within
    3e-4 of the minimum) must pick exactly what the exact sweep over all runs picks: same subset size, log p-value and PWM."""
    import os
    if not mc.have_reference():
        pytest.skip("oracle/_ref/ref_harness not built")
    w = W.by_name(name).scaled(nseq)
    ref, _, _, _ = mc.run_reference(W.by_name(name).scaled(1500), 1)  # only for the start motif
    n, codes = w.positive_set()
    ref = dict(ref, posSetSize=len(n), total=int(n.astype(np.int64).sum()))
    os.environ.pop("XXM_NO_SWEEP_PREFILTER", None)
    a = gpu_steps(w, ref, n, codes, None, 3)
    os.environ["XXM_NO_SWEEP_PREFILTER"] = "1"
    try:
        b = gpu_steps(w, ref, n, codes, None, 3)
    finally:
        os.environ.pop("XXM_NO_SWEEP_PREFILTER", None)
    for k in ("nbSites", "nbBestSites", "maxSites", "bestMotifNb", "bestPval", "enr"):
        assert a[k] == b[k], "%s: %r (prefilter) != %r (full sweep)" % (k, a[k], b[k])
    assert (np.asarray(a["pwm"]) == np.asarray(b["pwm"])).all()
    assert a["bestMotifNb"] > 100
