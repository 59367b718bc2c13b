"""Host-side pieces of the product (table construction, optimal-subset sweep) against the reference's golden vectors.
These run without a GPU through the device-free entry points of the C ABI."""
import numpy as np
import pytest

import xxmotif_b200.capi as capi
from backends import tables_dump_dense
from conftest import golden_names
from replay import TraceInputs, _cmp_tables, _eq


def _table_prefixes(t):
    out = [k[:-len(".tab.L")] for k in t if k.endswith(".tab.L")]
    return out


@pytest.mark.parametrize("name", golden_names())
def test_host_tables_vs_reference(golden, name):
    t = golden(name)
    if "nm.rprobs" not in t:
        pytest.skip("order-8 NullModel tables are not in the fixture")
    ti = TraceInputs(t)
    n = 0
    for p in _table_prefixes(t):
        if p.startswith("opt"):
            pwm, cols, pss = t[p + ".in.pwm"], t[p + ".in.cols"], int(t[p + ".in.posSetSize"][0])
        elif p.startswith("el"):
            pwm, cols, pss = t[p + ".pwm_in"], t[p + ".cols"], int(t["opt1.in.posSetSize"][0])
        else:
            pwm, cols, pss = t["converged.pwm"], t["converged.cols"], int(t["converged.posSetSize"][0])
        d = capi.host_tables(pwm, cols, pss, int(ti.total[pss]), ti.bgLog, ti.order, t["nm.rprobs"], t["nm.rconds"])
        _cmp_tables(p + ".tab", tables_dump_dense(d["L"], d["maxk"], d["nparts"], d["pvalues"], d["probs"], d["classes"]), t, p + ".tab")
        n += 1
    assert n >= 3


@pytest.mark.parametrize("name", golden_names())
def test_host_best_subset_vs_reference(golden, name):
    t = golden(name)
    ti = TraceInputs(t)
    nent = len(ti.n_post)
    checked = 0
    for p in ["opt0", "opt1"] + ["el%d.%d" % (v, it) for v in range(int(t["el.count"][0])) for it in range(3)]:
        if p + ".sorted.score" not in t:
            continue
        sc, sq = t[p + ".sorted.score"], t[p + ".sorted.seq"]
        nb_key = p + (".bestMotifNb" if p.startswith("opt") else ".motifNb")
        bp_key = p + (".bestPval" if p.startswith("opt") else ".overrepPval")
        if len(sc) == 0 or sc[0] == 1:
            continue
        pss = int(t["opt1.in.posSetSize"][0])
        nb, bp = capi.host_best_subset(sc, np.where(sc < 1, sq, 1), pss, int(t[p + ".maxSites"][0]), nent, oneOccurrence=bool(ti.oops))
        _eq(p + ".bestMotifNb", nb, int(t[nb_key][0]))
        _eq(p + ".bestPval", bp, float(t[bp_key][0]))
        checked += 1
    assert checked >= 2
