"""Row a15 (localization statistics): oracle and product host code against the reference's golden values (CPU), and the
device histogram path (GPU)."""
import numpy as np
import pytest

from conftest import golden_names
from replay import TraceInputs

LOC = [n for n in golden_names() if n.startswith("c5")]


def _best_sorted(t, p="opt1"):
    import oracle_lib as ol
    b = np.zeros(len(t[p + ".best.seq"]), dtype=ol.SITE_DTYPE)
    b["sequence"], b["startPos"], b["score"] = t[p + ".best.seq"], t[p + ".best.pos"], t[p + ".best.score"]
    return b


@pytest.mark.parametrize("name", LOC)
def test_oracle_enriched_region(golden, name):
    import oracle_region
    from backends import OracleBackend
    t = golden(name)
    ti = TraceInputs(t)
    be = OracleBackend(ti, from_scratch=False)
    cols = t["opt1.in.cols"]
    got = oracle_region.update_enriched_region(be.ctx, _best_sorted(t), ti.revcomp, int(t["in.pos.max_leng"][0]), int(cols[0]),
                                               int(cols[-1] - cols[0] + 1), t["opt1.in.enrich"])
    assert got == tuple(int(x) for x in t["optA.enriched.enrich"])


@pytest.mark.parametrize("name", golden_names())
def test_positional_pval_host(golden, name):
    """setPositionalPval of the converged motif (product host code and oracle) vs the reference's posPval.  The reference
    sums log(pValPos) over the site list in SCORE order (maximizeMotifLength, Iterate_Motif.cpp:146) and only afterwards
    sorts the list by position (MotifContainer.cpp:262); the trace keeps the sorted list, so the sum order differs and the
    comparison with the reference is to 1e-12 relative; product and oracle must agree exactly."""
    from backends import OracleBackend
    from xxmotif_b200 import localization
    t = golden(name)
    ti = TraceInputs(t)
    cols = t["converged.cols"]
    enr = t["converged.enrich"]
    want = float(t["converged.posPval"][0])
    got = localization.positional_pval(int(t["in.pos.max_leng"][0]), ti.revcomp, int(enr[0]), int(cols[0]), int(cols[-1] - cols[0] + 1),
                                       t["converged.sites.pos"])
    assert got == pytest.approx(want, rel=1e-12, abs=1e-300)
    be = OracleBackend(ti, from_scratch="nm.rprobs" not in t)
    assert be.ctx.positional_pval(int(enr[0]), int(cols[0]), int(cols[-1] - cols[0] + 1), t["converged.sites.pos"]) == got


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOC)
def test_gpu_enriched_region(golden, name):
    from gpu_backend import GpuBackend
    from xxmotif_b200 import localization
    t = golden(name)
    ti = TraceInputs(t)
    be = GpuBackend(ti)
    try:
        p = "opt1"
        pwm, cols = t[p + ".in.pwm"], t[p + ".in.cols"]
        pss = int(t[p + ".in.posSetSize"][0])
        e = t[p + ".in.enrich"]
        be.ctx.reinitialize(pwm, cols, pss, int(ti.total[pss]))
        be.ctx.scan_full(pss, (int(e[1]), int(e[2])))
        got = localization.update_enriched_region(be.ctx, int(cols[0]), int(cols[-1] - cols[0] + 1), e)
        assert got == tuple(int(x) for x in t["optA.enriched.enrich"])
    finally:
        be.ctx.close()
