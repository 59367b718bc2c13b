"""GPU parity, callers of the path: IterateMotif::optimizeMotif, ::maximizeMotifLength and the per-motif
loop of MotifContainer::iterate, re-stated in csrc/refine.cpp above the C ABI, against the reference's own black-box runs
of those functions recorded in the golden traces (ref_harness stages A, E, F).  Bit-exact: PWM, match columns, site list,
log p-values, enriched region."""
import numpy as np
import pytest

from conftest import golden_names
from replay import TraceInputs

pytestmark = pytest.mark.gpu


def _cmp_motif(tag, m, t, p, sites_sorted=False):
    assert np.array_equal(m.cols, t[p + ".cols"]), (tag, "cols", m.cols, t[p + ".cols"])
    a, b = m.pwm.reshape(31, 5), np.asarray(t[p + ".pwm"]).reshape(31, 5)
    for c in m.cols:  # columns outside the motif hold leftovers of earlier variants in both implementations; only the motif's are defined
        assert np.array_equal(a[c, 1:], b[c, 1:]), (tag, "pwm column", c, a[c], b[c])
    assert np.array_equal(m.enr, t[p + ".enrich"]), (tag, "enrich", m.enr, t[p + ".enrich"])
    assert m.pval == float(t[p + ".pval"][0]), (tag, "pval", m.pval, float(t[p + ".pval"][0]))
    assert m.posPval == float(t[p + ".posPval"][0]), (tag, "posPval", m.posPval, float(t[p + ".posPval"][0]))
    assert np.array_equal(m.siteSeq, t[p + ".sites.seq"]), (tag, "sites.seq")
    assert np.array_equal(m.sitePos, t[p + ".sites.pos"]), (tag, "sites.pos")


@pytest.mark.parametrize("name", golden_names())
def test_refine_callers_gpu(golden, name):
    from gpu_backend import GpuBackend
    from xxmotif_b200 import localization, refine
    t = golden(name)
    ti = TraceInputs(t)
    be = GpuBackend(ti)
    try:
        pss = int(t["m0.posSetSize"][0])
        m = refine.Motif(t["m0.pwm"], t["m0.cols"], pss, enr=t["m0.enrich"], siteSeq=t["m0.sites.seq"], sitePos=t["m0.sites.pos"],
                         pval=float(t["m0.pval"][0]), posPval=float(t["m0.posPval"][0]))
        R = refine.Refiner(be.ctx, totalPositions=int(ti.total[pss]), revcomp=ti.revcomp, localization=ti.localization,
                           maximizeMotifLength=int(t["cfg.maximizeMotifLength"][0]), minMatchPositions=int(t["cfg.minMatchPositions"][0]),
                           maxMatchPositions=int(t["cfg.maxMatchPositions"][0]), pseudo=ti.pseudo, neff_pwm=float(t["cfg.neff_pwm"][0]))
        # stage A: optimizeMotif
        R.optimize_motif(m)
        assert np.array_equal(m.pwm[5:], t["opt1.pwm_out"][5:])
        assert m.posPval == float(t["optA.posPval"][0])
        if ti.localization:
            cols = m.cols
            e = localization.update_enriched_region(be.ctx, int(cols[0]), int(cols[-1] - cols[0] + 1), m.enr)
            for i in range(3):
                m.m.enr[i] = e[i]
            assert np.array_equal(m.enr, t["optA.enriched.enrich"])
        # stage E: maximizeMotifLength
        R.maximize_motif_length(m)
        _cmp_motif("maxlen", m, t, "maxlen")
        # stage F: iterate to convergence
        rounds = R.iterate_motif(m)
        assert rounds >= 1
        _cmp_motif("converged", m, t, "converged")
    finally:
        be.ctx.close()
