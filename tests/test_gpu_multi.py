"""xxm_multi (include/xxmotif_gpu.h): one host thread driving several GPUs.  Every call the reference-side binding makes
(reinitialize, full scan, candidate re-score, OOPS scan, start-position-list re-score, the lists) must give, over shards of the set
on N devices, exactly what one context gives over the whole set: counts, maxSites, and the lists element for element (scores
bit for bit -- the shards work with the tables of the whole set).  N = 1 (pass-through) runs on any box; N = 2 and N = 8 need as many GPUs
(gpurun --gpus N)."""
import numpy as np
import pytest

import matrix_common as mc
from xxmotif_b200 import capi
from xxmotif_b200 import workloads as W

pytestmark = pytest.mark.gpu

CASES = {
    "zoops": W.Workload("a", 6000, 300, 41, frac=0.6),
    "mops_revcomp": W.Workload("b", 5000, 400, 42, revcomp=True, mops=True),
    "ragged_12_columns": W.Workload("c", 6000, 0, 43, ragged=(60, 900), start=("p", 12)),
    "localization_gapped": W.Workload("d", 6000, 300, 44, frac=0.6, planted="TATAW.AWR", start=("m", "TATAW.AWR"), localization=True, anchor=170, downstream=100),
    "order4_negset": W.Workload("e", 4000, 200, 45, revcomp=True, order=4, neg=9000),
}


def _calls(ctx, w, ref, n, codes, neg, oops):
    bg, bgl, rp, rc = mc.background_inputs(w, n, codes, neg)
    ctx.set_options(mops=w.mops, oops=oops, localization=w.localization, bg=bg, bgLog=bgl)
    ctx.set_background(w.order, rp, rc)
    ctx.load_seqset(n, codes, revcomp=w.revcomp)
    nent = len(n)
    total = int(ref["total"])
    cols = np.asarray(ref["cols"], dtype=np.int32)
    pwm = np.asarray(ref["pwm0"], dtype=np.float64)
    enr = (ref["enr0"][1], ref["enr0"][2])
    out = {}
    ctx.reinitialize(pwm, cols, nent, total)
    if oops:
        out["oops"] = ctx.scan_oops(nent, enr)
        out["oops_sites"] = ctx.get_sites(capi.LIST_SORTED_SITES)
        return out
    out["scan"] = ctx.scan_full(nent, enr)
    out["best_by_score"] = ctx.get_sites(capi.LIST_BEST_STARTPOS, capi.ORDER_BY_SCORE)
    out["sorted"] = ctx.get_sites(capi.LIST_SORTED_SITES)
    # the elongation phase's re-score: the model one column further right, the candidates of the scan
    if cols[-1] < 30:
        cols2 = np.append(cols, cols[-1] + 1).astype(np.int32)
        pwm2 = pwm.reshape(31, 5).copy()
        pwm2[cols2[-1]] = [0.0] + list(np.log([0.125, 0.25, 0.125, 0.5]))
        ctx.reinitialize(pwm2, cols2, nent, total)
        out["rescore"] = ctx.rescore(int(cols[0]), nent, enr)
        out["elong"] = ctx.get_sites(capi.LIST_SORTED_SITES_ELONG)
    # a start position list re-scored under the first model (every third kept site)
    ctx.reinitialize(pwm, cols, nent, total)
    pick = out["sorted"][::3]
    pick = pick[pick["startPos"] > 0]
    out["spl"] = ctx.rescore_startposlist(nent, enr, pick["sequence"].astype(np.int32), pick["startPos"].astype(np.int32), instanceThreshold=False)
    out["spl_sites"] = ctx.get_sites(capi.LIST_SORTED_SITES)
    return out


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert len(a[k]) == len(b[k]), k
            for f in ("sequence", "startPos"):
                assert np.array_equal(a[k][f], b[k][f]), (k, f)
            assert np.array_equal(a[k]["score"].view(np.uint32), b[k]["score"].view(np.uint32)), (k, "score")
        else:
            assert a[k] == b[k], k


@pytest.mark.parametrize("ngpus", [1, 2, 8])
@pytest.mark.parametrize("oops", [False, True], ids=["zoops_mops", "oops"])
@pytest.mark.parametrize("name", list(CASES))
def test_multi_device_context_gives_the_single_context_results(name, oops, ngpus):
    import torch
    if torch.cuda.device_count() < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    if not mc.have_reference():
        pytest.skip("oracle/_ref/ref_harness not built (start PWM comes from the reference)")
    w = CASES[name]
    if oops and w.mops:
        pytest.skip("--oops excludes --mops")
    ref, n, codes, neg = mc.run_reference(w, 1)
    one = capi.Context(0)
    many = capi.MultiContext(ngpus)
    try:
        a = _calls(one, w, ref, n, codes, neg, oops)
        b = _calls(many, w, ref, n, codes, neg, oops)
        assert len(a.get("sorted", a.get("oops_sites"))) > 100
        _same(a, b)
    finally:
        one.close()
        many.close()


def test_multi_device_context_refuses_partial_sets():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = CASES["zoops"]
    n, codes = w.positive_set()
    many = capi.MultiContext(2)
    try:
        bg, bgl, rp, rc = mc.background_inputs(w, n, codes, None)
        many.set_options(bg=bg, bgLog=bgl)
        many.set_background(w.order, rp, rc)
        many.load_seqset(n, codes)
        with pytest.raises(capi.XxmError, match="whole set"):
            many.scan_full(len(n) - 1, (1, 1))
    finally:
        many.close()
