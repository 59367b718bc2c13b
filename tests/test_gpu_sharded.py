"""GPU: the sharded step -- split-phase entry points with the exchange in Python, and the one-call native step
(xxm_optimize_step) -- equals the reference's update_PWM_with_bestSubset on the golden traces."""
import numpy as np
import pytest

from conftest import golden_names
from replay import TraceInputs, _eq

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("name", ["c1_zoops", "c2_mops_rc", "c6_long12_mops", "c8_ragged_n", "c3_neg_o4_rc"])
def test_split_phase_step_matches_reference(golden, name, native):
    from gpu_backend import GpuBackend
    from xxmotif_b200 import sharded
    t = golden(name)
    ti = TraceInputs(t)
    be = GpuBackend(ti)
    try:
        for it in range(2):
            p = "opt%d" % it
            pwm, cols = t[p + ".in.pwm"], t[p + ".in.cols"]
            pss = int(t[p + ".in.posSetSize"][0])
            e = t[p + ".in.enrich"]
            job = sharded.ShardedRefinement(be.ctx, None, nent_local=pss, total_local=int(ti.total[pss]), pseudo=ti.pseudo, revcomp=bool(ti.revcomp),
                                            oneOccurrence=bool(ti.oops), bg=ti.bg, native=native)
            new, info = job.optimize_step(pwm, cols, (int(e[1]), int(e[2])))
            _eq(p + ".nbSites", info["nbSites"], int(t[p + ".nbSites"][0]))
            _eq(p + ".maxSites", info["maxSites"], int(t[p + ".maxSites"][0]))
            _eq(p + ".bestMotifNb", info["bestMotifNb"], int(t[p + ".bestMotifNb"][0]))
            _eq(p + ".bestPval", info["bestPval"], float(t[p + ".bestPval"][0]))
            _eq(p + ".pwm_out", new[5:], t[p + ".pwm_out"][5:])
    finally:
        be.ctx.close()
