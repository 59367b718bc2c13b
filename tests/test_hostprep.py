"""Host-side input preparation of the product (revcomp doubling, background, NullModel) vs the reference's values."""
import numpy as np
import pytest

from conftest import golden_names
from replay import TraceInputs, _eq
from xxmotif_b200 import hostprep


@pytest.mark.parametrize("name", golden_names())
def test_background_and_nullmodel(golden, name):
    t = golden(name)
    ti = TraceInputs(t)
    n_pre, S_pre = ti.pos_pre()
    if ti.revcomp:
        n2, S2 = hostprep.revcomp_double(n_pre, S_pre)
        _eq("revcomp.n", n2, ti.n_post)
        _eq("revcomp.S", S2, ti.S_post)
    bg, bgl = hostprep.background(ti.n_post, ti.S_post)
    _eq("posBg", bg[1:], t["bg.posBg"][1:])
    _eq("posBg_log", bgl[1:], t["bg.posBg_log"][1:])
    if ti.hasNeg:
        n, S = t["in.neg.n"], t["in.neg.S"]
        nbg, nbgl = hostprep.background(n, S)
        _eq("negBg", nbg[1:], t["bg.negBg"][1:])
    else:
        n, S = ti.n_post, ti.S_post
    if "nm.rprobs" in t:
        rp, rc = hostprep.null_model(n, S, order=ti.order, alpha=float(t["cfg.pseudocountsFactor"][0]), beta=float(t["cfg.countsOffset"][0]))
        _eq("rprobs", rp, t["nm.rprobs"])
        _eq("rconds", rc, t["nm.rconds"])


def test_profile_pwm_matches_reference(golden):
    t = golden("c4_prof")
    pwm, cols = hostprep.profile_pwm(hostprep.TATA_PROFILE / 100.0)  # the fixture's profile file holds fractions
    _eq("cols", cols, t["m0.cols"])
    _eq("pwm", pwm[5:], t["m0.pwm"][5:])


def test_synthetic_set_shape():
    n, codes = hostprep.synthetic_set(100, 50, seed=1)
    assert n.sum() == codes.size and codes.min() >= 1 and codes.max() <= 4
