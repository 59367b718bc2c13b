"""the p-value tables of one-part models built on the DEVICE (xxm_device_tables: parallel test of all
branch-and-bound paths, rank-by-counting stable sort, sequential class accumulation) against the tables of the UNMODIFIED
reference in every golden trace (ids, p-values bit for bit, classes with their id lists in the reference's order), against the
host build on random PWMs and gapped column sets, and as the table source of a whole replay."""
import os

import numpy as np
import pytest

from conftest import golden_names
from replay import TraceInputs, _cmp_tables, replay
from backends import tables_dump_dense

pytestmark = pytest.mark.gpu


def _dump(d):
    return tables_dump_dense(d["L"], d["maxk"], d["nparts"], d["pvalues"], d["probs"], d["classes"])


@pytest.mark.parametrize("name", golden_names())
def test_device_tables_vs_reference(golden, name):
    from gpu_backend import GpuBackend
    t = golden(name)
    ti = TraceInputs(t)
    tabs = [k[:-len(".tab.L")] for k in t if k.endswith(".tab.L") and int(t[k[:-2] + ".hashNumber"][0]) == 1]
    if not tabs:
        pytest.skip("no one-part table in this trace")
    be = GpuBackend(ti)
    try:
        n = 0
        for p in tabs:
            pre = p if (p + ".in.pwm") in t else None
            if p.startswith("opt"):
                pwm, cols, pss = t[p + ".in.pwm"], t[p + ".in.cols"], int(t[p + ".in.posSetSize"][0])
            elif p.startswith("el"):
                pwm, cols, pss = t[p + ".pwm_in"], t[p + ".cols"], int(t["opt0.in.posSetSize"][0])
            else:
                continue
            d, ms = be.ctx.device_tables(pwm, cols, pss, int(ti.total[pss]))
            _cmp_tables(name + ":" + p, _dump(d), t, p + ".tab")
            n += 1
        assert n > 0
    finally:
        be.ctx.close()


def test_device_tables_vs_host_random():
    """random PWMs, 5 to 8 match columns with and without gap columns, background orders 0..4: device == host build"""
    import xxmotif_b200 as xb
    from xxmotif_b200 import capi, hostprep
    rng = np.random.default_rng(12)
    n, codes = hostprep.synthetic_set(300, 200, seed=3)
    ctx = xb.Context(0)
    try:
        for trial in range(24):
            order = int(rng.integers(0, 5))
            bg, bgl = hostprep.background(n, codes)
            rp, rc = hostprep.null_model(n, codes, order=order)
            ctx.set_options(bg=bg, bgLog=bgl)
            ctx.set_background(order, rp, rc)
            L = int(rng.integers(5, 9))
            first = int(rng.integers(1, 8))
            gaps = rng.integers(0, 3, size=L - 1) * (rng.random(L - 1) < 0.3)
            cols = (first + np.concatenate(([0], np.cumsum(1 + gaps)))).astype(np.int32)
            pwm = np.zeros((31, 5))
            pwm[1:, 1:] = np.log(0.25)
            for c in cols:
                pr = rng.dirichlet(np.ones(4) * rng.choice([0.15, 0.5, 2.0]))
                pr = np.maximum(pr, 1e-3)
                pr /= pr.sum()
                pwm[c, 1:] = np.log(pr)
            want = capi.host_tables(pwm.reshape(-1), cols, 300, 60000, bgl, order, rp, rc)
            got, ms = ctx.device_tables(pwm.reshape(-1), cols, 300, 60000)
            a, b = _dump(got), _dump(want)
            for k in ("ids", "vals", "cscore", "ccount", "cids"):
                assert (a[k][0] == b[k][0]).all(), "trial %d (%d columns, order %d): %s differs" % (trial, L, order, k)
            assert a["L"] == b["L"] and a["nparts"] == 1
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["c1_zoops", "c2_mops_rc", "c5_loc_gap", "c7_oops", "c11_short5"])
def test_replay_with_device_tables(golden, name):
    """every stage of a trace with xxm_reinitialize taking its one-part tables from the device build"""
    from gpu_backend import GpuBackend
    t = golden(name)
    os.environ["XXM_DEVICE_TABLES"] = "1"
    try:
        be = GpuBackend(TraceInputs(t))
        try:
            done = replay(t, be)
            assert "opt1" in done
        finally:
            be.ctx.close()
    finally:
        os.environ.pop("XXM_DEVICE_TABLES", None)
