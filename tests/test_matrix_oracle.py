"""The oracle (C restatement, oracle/xxm_oracle.c) against the UNMODIFIED reference at a few thousand sequences, for every
shape of the workload matrix (BASELINE.json configs 2-5, 12 and 17 match columns, ragged lengths): three optimizeMotif steps
end on the same site counts, subset choice, p-value and PWM -- bit for bit.  The golden traces pin the oracle stage by stage on
40-80 sequences; this pins it end to end at a size where score ties, cluster replacement and both table parts occur in bulk."""
import numpy as np
import pytest

import matrix_common as mc
from xxmotif_b200 import workloads as W

STEPS = 3
SMALL = {"cfg2": 2500, "cfg3": 3000, "cfg4": 1500, "cfg5": 4000, "cols12": 1500, "cols17": 600, "ragged": 1500}


@pytest.mark.parametrize("name", [w.name for w, _ in W.MATRIX])
def test_oracle_steps_match_reference(name):
    if not mc.have_reference():
        pytest.skip("oracle/_ref/ref_harness not built (make -C oracle ref)")
    import oracle_lib as ol
    import oracle_region
    w = W.by_name(name).scaled(SMALL[name], neg=4000 if W.by_name(name).neg else None)
    ref, n, codes, neg = mc.run_reference(w, STEPS)
    bg, bgl, rp, rc = mc.background_inputs(w, n, codes, neg)
    oset = ol.SeqSet(n, codes, revcomp=w.revcomp)
    onm = ol.NullModel(order=w.order, rprobs=rp, rconds=rc)
    octx = ol.Ctx(oset, w.revcomp, w.mops, False, w.localization, bg, bgl)
    nent = len(n)
    assert (nent, int(oset.total[nent])) == (ref["posSetSize"], ref["total"])
    cols = np.asarray(ref["cols"], dtype=np.int32)
    span = int(cols[-1] - cols[0] + 1)
    pwm = np.asarray(ref["pwm0"], dtype=np.float64)
    enr = tuple(ref["enr0"])
    got = {}
    for _ in range(STEPS):
        tab = ol.Tables(pwm, cols, nent, int(oset.total[nent]), bgl, onm)
        assert tab.error == 0
        octx.scan_full(tab, cols, nent, (enr[1], enr[2]), True)
        got.update(nbSites=octx.nbSites, nbBestSites=octx.nbBestSites, maxSites=octx.maxSites)
        best = ol.sort_by_score(octx.best()) if w.localization else None
        ok, nb, bp, new = octx.update_pwm_with_best_subset(0, pwm, cols, nent, 0, False, 0.1)
        if ok:
            pwm = new
        got.update(bestMotifNb=nb, bestPval=bp)
        if w.localization:
            enr = oracle_region.update_enriched_region(octx, best, w.revcomp, oset.max_leng, int(cols[0]), span, enr)
    got["pwm"] = pwm
    got["enr"] = enr
    mc.compare_with_reference(ref, got)
