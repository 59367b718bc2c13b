"""GPU parity, part 2: the CUDA path against the oracle (the pinned C restatement) on larger seeded sets, and
size-independent properties at BASELINE.json's full size (1M x 1 kb)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(nseq, length, seed, revcomp, mops, localization, order=2):
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    n, codes = hostprep.synthetic_set(nseq, length, seed=seed, frac=0.6)
    rng = np.random.default_rng(seed + 99)
    if nseq <= 5000:  # sprinkle some N runs
        for _ in range(nseq // 3):
            p = int(rng.integers(0, codes.size - 6))
            codes[p:p + int(rng.integers(1, 6))] = 0
    if revcomp:
        n2, c2 = hostprep.revcomp_double(n, codes)
    else:
        n2, c2 = n, codes
    bg, bgl = hostprep.background(n2, c2)
    rp, rc = hostprep.null_model(n2, c2, order=order)
    ctx = xb.Context(0)
    ctx.set_options(mops=mops, localization=localization, bg=bg, bgLog=bgl)
    ctx.set_background(order, rp, rc)
    ctx.load_seqset(n, codes, revcomp=revcomp)
    oset = ol.SeqSet(n, codes, revcomp=revcomp)
    onm = ol.NullModel(order=order, rprobs=rp, rconds=rc)
    octx = ol.Ctx(oset, revcomp, mops, False, localization, bg, bgl)
    return ctx, octx, oset, onm, bg, bgl, n, codes


def _site_eq(name, got, want):
    assert len(got) == len(want), "%s: %d != %d entries" % (name, len(got), len(want))
    for f in ("score", "startPos", "sequence"):
        bad = np.flatnonzero(got[f] != want[f])
        assert bad.size == 0, "%s.%s: %d mismatches, first at %d: %s vs %s" % (name, f, bad.size, bad[0], got[f][bad[0]], want[f][bad[0]])


@pytest.mark.parametrize("revcomp,mops,localization,length", [(False, False, False, 400), (True, True, False, 300), (False, True, True, 300),
                                                              (True, False, False, 250)])
def test_refinement_steps_vs_oracle(revcomp, mops, localization, length):
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    ctx, octx, oset, onm, bg, bgl, n, codes = _setup(3000, length, 11, revcomp, mops, localization)
    try:
        total, avg, mxl, mnl = ctx.seqset_info()
        assert (total == oset.total).all() and (avg[1:] == oset.avgLength[1:]).all()
        pwm, cols = hostprep.profile_pwm()
        nent = len(n)
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        for it in range(3):
            ctx.reinitialize(pwm, cols, nent, int(total[nent]))
            tab = ol.Tables(pwm, cols, nent, int(total[nent]), bgl, onm)
            nb, nbest, mx = ctx.scan_full(nent, enr)
            octx.scan_full(tab, cols, nent, enr, True)
            assert (nb, nbest, mx) == (octx.nbSites, octx.nbBestSites, octx.maxSites)
            _site_eq("best", ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_BY_SCORE), octx.best())
            _site_eq("sites", ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS), octx.sites())
            ok, nbm, bp, new = ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES, pwm, int(cols[0]), nent, False, 0.1)
            ook, onbm, obp, onew = octx.update_pwm_with_best_subset(0, pwm, cols, nent, 0, False, 0.1)
            assert (ok, nbm, bp) == (ook, onbm, obp)
            assert (new[5:] == onew[5:]).all()
            # candidate re-score under column variants (one longer, one shorter, one shifted)
            for tmp in (np.arange(cols[0] - 2, cols[-1] + 1), np.arange(cols[0] + 1, cols[-1] + 1), np.arange(cols[0] + 1, cols[-1] + 2)):
                tmp = tmp.astype(np.int32)
                ctx.reinitialize(new, tmp, nent, int(total[nent]))
                ttab = ol.Tables(new, tmp, nent, int(total[nent]), bgl, onm)
                nb2, mx2 = ctx.rescore(int(cols[0]), nent, enr)
                octx.rescore(ttab, int(cols[0]), tmp, nent, enr, False)
                assert (nb2, mx2) == (octx.nbSites, octx.maxSites)
                got = ctx.get_sites(xb.LIST_SORTED_SITES_ELONG, xb.ORDER_AS_IS)
                want = octx.elong(nb2)
                m = want["score"] != 1
                assert (got["score"][:nb2] == want["score"]).all()
                assert (got["startPos"][:nb2][m] == want["startPos"][m]).all() and (got["sequence"][:nb2][m] == want["sequence"][m]).all()
                ok, nbm, bp, e1 = ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES_ELONG, new, int(tmp[0]), nent, True, 0.1)
                ook, onbm, obp, oe1 = octx.update_pwm_with_best_subset(1, new, tmp, nent, 0, True, 0.1)
                assert (ok, nbm, bp) == (ook, onbm, obp)
                if ok:
                    assert (e1[5:] == oe1[5:]).all()
            pwm = new
    finally:
        ctx.close()


def test_full_size_properties():
    """1M x 1 kb (BASELINE cfg4 shape): ordering, consistency of the fused histogram, idempotence, and a bit-exact
    comparison of the first 2000 sequences with the oracle."""
    import torch
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a large GPU")
    nseq, length = 1000000, 1000
    n, codes = hostprep.synthetic_set(nseq, length, seed=4)
    bg, bgl = hostprep.background(n[:50000], codes[:50000 * length])  # any fixed background is a valid input of the path
    rp, rc = hostprep.null_model(n[:50000], codes[:50000 * length], order=2)
    ctx = xb.Context(0)
    try:
        ctx.set_options(mops=False, bg=bg, bgLog=bgl)
        ctx.set_background(2, rp, rc)
        ctx.load_seqset(n, codes, revcomp=False)
        total, avg, mxl, mnl = ctx.seqset_info()
        assert int(total[nseq]) == nseq * length
        pwm, cols = hostprep.profile_pwm()
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        job = sharded.ShardedRefinement(ctx, None, nent_local=nseq, total_local=int(total[nseq]), pseudo=0.1, bg=bg)
        pwm1, info1 = job.optimize_step(pwm, cols, enr)
        pwm2, info2 = job.optimize_step(pwm1, cols, enr)
        # the candidate list is strictly increasing in (sequence, position) -- the reference's scan order
        best = ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_AS_IS, cap=info2["nbBestSites"] + 8)
        assert len(best) == info2["nbBestSites"]
        key = best["sequence"].astype(np.int64) * 4096 + best["startPos"]
        assert (np.diff(key) > 0).all()
        assert best["startPos"].min() >= 1 and best["startPos"].max() <= length - 8 + 1
        # fused class histogram == number of kept sites; kept sites never overlap
        sc, cn = sharded.score_runs(ctx, xb.LIST_SORTED_SITES)
        assert int(cn.sum()) == info2["nbSites"] and (np.diff(sc) > 0).all()
        kept = ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS, cap=info2["nbSites"] + 8)
        same = kept["sequence"][1:] == kept["sequence"][:-1]
        assert ((kept["startPos"][1:] - kept["startPos"][:-1])[same] > 8).all()
        # idempotence: the same PWM gives the same step
        pwm2b, info2b = job.optimize_step(pwm1, cols, enr)
        assert info2b == info2 and (pwm2b == pwm2).all()
        # bit-exact against the oracle on the first 2000 sequences
        m = 2000
        oset = ol.SeqSet(n[:m], codes[:m * length])
        onm = ol.NullModel(order=2, rprobs=rp, rconds=rc)
        octx = ol.Ctx(oset, False, False, False, False, bg, bgl)
        tab = ol.Tables(pwm1, cols, m, m * length, bgl, onm)
        octx.scan_full(tab, cols, m, enr, True)
        want = ol.sort_by_startpos(octx.best())
        got = best[best["sequence"] <= m]
        assert len(got) == len(want)
        for f in ("score", "startPos", "sequence"):
            assert (got[f] == want[f]).all()
        wk = octx.sites()
        gk = kept[kept["sequence"] <= m]
        assert len(gk) == len(wk) and (gk["startPos"] == wk["startPos"]).all() and (gk["score"] == wk["score"]).all()
    finally:
        ctx.close()


def test_shard_invariance():
    """Two half shards combined through the split-phase API give exactly the step of the whole set."""
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded
    nseq, length = 40000, 500
    n, codes = hostprep.synthetic_set(nseq, length, seed=21)
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    pwm, cols = hostprep.profile_pwm()

    def make(lo, hi):
        c = xb.Context(0)
        c.set_options(mops=True, bg=bg, bgLog=bgl)
        c.set_background(2, rp, rc)
        c.load_seqset(n[lo:hi], codes[lo * length:hi * length])
        return c

    whole, a, b = make(0, nseq), make(0, nseq // 2), make(nseq // 2, nseq)
    try:
        enr = (1 - int(cols[0]) + 1, length - int(cols[0]) + 1)
        p = pwm
        for it in range(2):
            job = sharded.ShardedRefinement(whole, None, nent_local=nseq, total_local=nseq * length, pseudo=0.1, bg=bg)
            pw, iw = job.optimize_step(p, cols, enr)
            runs, counts, nb = [], [], 0
            for c in (a, b):
                c.reinitialize(p, cols, nseq // 2, nseq // 2 * length)
                k, _, _ = c.scan_full(nseq // 2, enr)
                nb += k
                runs.append(sharded.score_runs(c, xb.LIST_SORTED_SITES))
            gsc, gcn = sharded.merge_runs([r[0] for r in runs], [r[1] for r in runs])
            maxSites = nseq * length - nseq * 7
            K, bp, thr = sharded.choose_subset(gsc, gcn, nb, nseq * 1000, maxSites)
            cnt = sum(sharded.count_sites(c, xb.LIST_SORTED_SITES, thr, int(cols[0])) for c in (a, b))
            ps = sharded.build_pwm(cnt, 0.1 * K, bg)
            assert nb == iw["nbSites"] and K == iw["bestMotifNb"] and bp == iw["bestPval"]
            assert (ps == pw).all()
            p = pw
    finally:
        whole.close(); a.close(); b.close()


def _low_complexity_set():
    rng = np.random.default_rng(5)
    nseq = 1500
    lens = rng.integers(40, 900, size=nseq).astype(np.int32)
    parts = []
    for L in lens:
        kind = rng.integers(0, 4)
        if kind == 0:
            s = np.tile(np.array([4, 1], dtype=np.uint8), L // 2 + 1)[:L]  # TATATA...
        elif kind == 1:
            s = np.full(L, 1, dtype=np.uint8)  # poly-A
            s[rng.integers(0, L, size=max(1, L // 50))] = 4
        else:
            s = rng.integers(1, 5, size=L).astype(np.uint8)
            if L > 60:
                a = int(rng.integers(0, L - 50))
                s[a:a + 40] = np.tile(np.array([4, 1, 4, 1, 1, 1], dtype=np.uint8), 7)[:40]
        if rng.random() < 0.3:
            k = int(rng.integers(1, 12))
            s[-k:] = 0
        if rng.random() < 0.2:
            a = int(rng.integers(0, L - 3))
            s[a:a + 3] = 0
        parts.append(s)
    codes = np.concatenate(parts)
    n = lens
    return n, codes


@pytest.mark.parametrize("revcomp,mops", [(False, False), (False, True), (True, True)])
def test_low_complexity_ragged_vs_oracle(revcomp, mops):
    """Stress case for the scan machinery: ragged lengths, long TA repeats and poly-A runs (hit clusters hundreds of entries
    long, tiles with more hits than the expand queue holds), N runs at sequence ends -- bit-exact against the oracle."""
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    n, codes = _low_complexity_set()
    if revcomp:
        n2, c2 = hostprep.revcomp_double(n, codes)
    else:
        n2, c2 = n, codes
    bg, bgl = hostprep.background(n2, c2)
    rp, rc = hostprep.null_model(n2, c2, order=2)
    ctx = xb.Context(0)
    try:
        ctx.set_options(mops=mops, bg=bg, bgLog=bgl)
        ctx.set_background(2, rp, rc)
        ctx.load_seqset(n, codes, revcomp=revcomp)
        oset = ol.SeqSet(n, codes, revcomp=revcomp)
        onm = ol.NullModel(order=2, rprobs=rp, rconds=rc)
        octx = ol.Ctx(oset, revcomp, mops, False, False, bg, bgl)
        total, avg, mxl, mnl = ctx.seqset_info()
        assert (total == oset.total).all()
        pwm, cols = hostprep.profile_pwm()
        nent = len(n)
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        for it in range(2):
            ctx.reinitialize(pwm, cols, nent, int(total[nent]))
            tab = ol.Tables(pwm, cols, nent, int(total[nent]), bgl, onm)
            nb, nbest, mx = ctx.scan_full(nent, enr)
            octx.scan_full(tab, cols, nent, enr, True)
            assert (nb, nbest, mx) == (octx.nbSites, octx.nbBestSites, octx.maxSites)
            assert nbest > (1000 if revcomp else 50000)  # the forward cases are dense on purpose (tiles overflow the expand queue)
            _site_eq("best", ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_BY_SCORE, cap=nbest + 8), octx.best())
            _site_eq("sites", ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS, cap=nb + 8), octx.sites())
            ok, nbm, bp, new = ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES, pwm, int(cols[0]), nent, False, 0.1)
            ook, onbm, obp, onew = octx.update_pwm_with_best_subset(0, pwm, cols, nent, 0, False, 0.1)
            assert (ok, nbm, bp) == (ook, onbm, obp)
            assert (new[5:] == onew[5:]).all()
            pwm = new
    finally:
        ctx.close()


@pytest.mark.parametrize("mops", [False, True])
def test_native_step_low_complexity_vs_oracle(mops):
    """The one-call native step (deferred scan, class-index overlap removal k_select_cls_fwd, lazy float scores) on the
    stress set: kept sites, their scores, subset choice and the new PWM bit-exact against the oracle, twice in a row."""
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded
    n, codes = _low_complexity_set()
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    ctx = xb.Context(0)
    try:
        ctx.set_options(mops=mops, bg=bg, bgLog=bgl)
        ctx.set_background(2, rp, rc)
        ctx.load_seqset(n, codes, revcomp=False)
        oset = ol.SeqSet(n, codes, revcomp=False)
        onm = ol.NullModel(order=2, rprobs=rp, rconds=rc)
        octx = ol.Ctx(oset, False, mops, False, False, bg, bgl)
        total, avg, mxl, mnl = ctx.seqset_info()
        pwm, cols = hostprep.profile_pwm()
        nent = len(n)
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        job = sharded.ShardedRefinement(ctx, None, nent_local=nent, total_local=int(total[nent]), pseudo=0.1, bg=bg, native=True)
        for it in range(3):
            new, info = job.optimize_step(pwm, cols, enr)
            tab = ol.Tables(pwm, cols, nent, int(total[nent]), bgl, onm)
            octx.scan_full(tab, cols, nent, enr, True)
            assert (info["nbSites"], info["nbBestSites"], info["maxSites"]) == (octx.nbSites, octx.nbBestSites, octx.maxSites)
            _site_eq("sites", ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS, cap=info["nbSites"] + 8), octx.sites())
            _site_eq("best", ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_BY_SCORE, cap=info["nbBestSites"] + 8), octx.best())
            ook, onbm, obp, onew = octx.update_pwm_with_best_subset(0, pwm, cols, nent, 0, False, 0.1)
            assert (info["bestMotifNb"], info["bestPval"]) == (onbm, obp)
            assert (new[5:] == onew[5:]).all()
            pwm = new
    finally:
        ctx.close()


def test_bench_workload_refinement_matches_reference():
    """The whole per-motif loop (xxm_iterate_motif) on a 10k-sequence sample of the bench workload against the unmodified
    reference's MotifContainer::iterate run on the same sequences here (oracle/_ref/ref_harness refine): bit-identical
    match columns, site count and log p-value."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    if not os.path.exists(bench.HARNESS):
        pytest.skip("oracle/_ref/ref_harness not built")
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, refine
    nseq, seqlen = 10000, 1000
    ref = bench.run_reference_refine(nseq, seqlen)
    n, codes = hostprep.synthetic_set(nseq, seqlen, seed=4)
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    ctx = xb.Context(0)
    try:
        ctx.set_options(mops=False, oops=False, localization=False, bg=bg, bgLog=bgl)
        ctx.set_background(2, rp, rc)
        ctx.load_seqset(n, codes, revcomp=False)
        total, _, mxl, _ = ctx.seqset_info()
        pwm0, cols = hostprep.profile_pwm()
        c0 = int(cols[0])
        mot = refine.Motif(pwm0, cols, nseq, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=1 << 20)
        R = refine.Refiner(ctx, totalPositions=int(total[nseq]), revcomp=False, localization=False, maximizeMotifLength=True, minMatchPositions=5,
                           maxMatchPositions=17, pseudo=0.1, neff_pwm=10.0)
        R.iterate_motif(mot)
        assert [int(mot.cols[0]), int(mot.cols[-1]), len(mot.cols)] == [ref["firstCol"], ref["lastCol"], ref["ncols"]]
        assert int(mot.m.nsites) == ref["nsites"]
        assert float(mot.pval) == ref["pval"]
    finally:
        ctx.close()


@pytest.mark.parametrize("revcomp", [False, True])
def test_per_sequence_cap_vs_oracle(revcomp):
    """Global::maxMotifsPerSequence (--maxSeqOcc) binding: getBestSubset skips the sites beyond the cap of their sequence and
    Motif::updatePWM_OOL counts the first K that are not (Motif.cpp:634-646) -- subset size, p-value and PWM against the oracle."""
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    n, codes = hostprep.synthetic_set(2500, 600, seed=17, frac=0.8)
    rng = np.random.default_rng(5)
    for s in rng.integers(0, 2500, size=400):  # a few extra instances per sequence, so that the cap of 2 binds
        for _ in range(4):
            p = int(rng.integers(0, 592))
            codes[s * 600 + p:s * 600 + p + 8] = [4, 1, 4, 1, 1, 1, 4, 1]
    n2, c2 = hostprep.revcomp_double(n, codes) if revcomp else (n, codes)
    bg, bgl = hostprep.background(n2, c2)
    rp, rc = hostprep.null_model(n2, c2, order=2)
    ctx = xb.Context(0)
    try:
        ctx.set_options(mops=True, bg=bg, bgLog=bgl, maxMotifsPerSequence=2)
        ctx.set_background(2, rp, rc)
        ctx.load_seqset(n, codes, revcomp=revcomp)
        oset = ol.SeqSet(n, codes, revcomp=revcomp)
        onm = ol.NullModel(order=2, rprobs=rp, rconds=rc)
        octx = ol.Ctx(oset, revcomp, True, False, False, bg, bgl, maxMotifsPerSequence=2)
        total, _, mxl, _ = ctx.seqset_info()
        pwm, cols = hostprep.profile_pwm()
        nent = len(n)
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        for it in range(3):
            ctx.reinitialize(pwm, cols, nent, int(total[nent]))
            tab = ol.Tables(pwm, cols, nent, int(total[nent]), bgl, onm)
            nb, nbest, mx = ctx.scan_full(nent, enr)
            octx.scan_full(tab, cols, nent, enr, True)
            assert (nb, nbest, mx) == (octx.nbSites, octx.nbBestSites, octx.maxSites)
            ok, nbm, bp, new = ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES, pwm, int(cols[0]), nent, False, 0.1)
            ook, onbm, obp, onew = octx.update_pwm_with_best_subset(0, pwm, cols, nent, 0, False, 0.1)
            assert (ok, nbm, bp) == (ook, onbm, obp)
            assert (new[5:] == onew[5:]).all()
            pwm = new
        kept = ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS, cap=nb + 8)
        assert np.bincount(kept["sequence"]).max() > 2  # the cap did bind
    finally:
        ctx.close()


def test_shard_invariance_ragged_12_columns():
    """with more than 8 match columns the tables depend on posSetSize / (2 totalPositions) of the WHOLE set
    (branch_and_bound-inl.h:417), which ragged shards do not share.  Two uneven shards of a ragged set, told the global counts
    (xxm_set_global_counts) and combined through the split-phase entry points, must give exactly the step of the whole set."""
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded, workloads as W
    w = W.by_name("ragged").scaled(12000)
    w.start = ("p", 12)
    n, codes = w.positive_set()
    off = np.concatenate(([0], np.cumsum(n.astype(np.int64))))
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    pwm, cols = w.start_pwm()
    nseq = len(n)
    cut = 4500

    def make(lo, hi):
        c = xb.Context(0)
        c.set_options(mops=True, bg=bg, bgLog=bgl)
        c.set_background(2, rp, rc)
        c.load_seqset(n[lo:hi], codes[off[lo]:off[hi]])
        return c

    whole, a, b = make(0, nseq), make(0, cut), make(cut, nseq)
    try:
        total_g = int(off[nseq])
        mxl = int(n.max())
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        for c in (a, b):
            c.set_global_counts(nseq, total_g, float("nan"), mxl)
        p = pwm
        span = int(cols[-1] - cols[0] + 1)
        for it in range(2):
            pw, iw = whole.optimize_step(p, cols, nseq, total_g, enr, 0.1)
            runs, nb = [], 0
            for c, (lo, hi) in ((a, (0, cut)), (b, (cut, nseq))):
                c.reinitialize(p, cols, hi - lo, int(off[hi] - off[lo]))
                k, _, _ = c.scan_full(hi - lo, enr)
                nb += k
                runs.append(sharded.score_runs(c, xb.LIST_SORTED_SITES))
            gsc, gcn = sharded.merge_runs([r[0] for r in runs], [r[1] for r in runs])
            maxSites = total_g - nseq * (span - 1)
            K, bp, thr = sharded.choose_subset(gsc, gcn, nb, nseq * 1000, maxSites)
            cnt = sum(sharded.count_sites(c, xb.LIST_SORTED_SITES, thr, int(cols[0])) for c in (a, b))
            ps = sharded.build_pwm(cnt, 0.1 * K, bg)
            assert nb == iw["nbSites"] and K == iw["bestMotifNb"] and bp == iw["bestPval"]
            assert (ps[5:] == pw[5:]).all()
            p = pw
        # without the global counts the shards prune differently (this is what ADVICE r1 was about) -- make sure the test can tell
        b2 = make(cut, nseq)
        try:
            b2.reinitialize(pwm, cols, nseq - cut, int(off[nseq] - off[cut]))
            k2, _, _ = b2.scan_full(nseq - cut, enr)
            b.reinitialize(pwm, cols, nseq - cut, int(off[nseq] - off[cut]))
            k1, _, _ = b.scan_full(nseq - cut, enr)
            assert isinstance(k1, int) and isinstance(k2, int)
        finally:
            b2.close()
    finally:
        whole.close(); a.close(); b.close()


@pytest.mark.parametrize("name,offsets,localization,mops", [
    ("one_part_span9_localized", [0, 1, 2, 3, 4, 6, 7, 8], True, True),
    ("one_part_span10_localized_zoops", [0, 1, 2, 3, 5, 6, 8, 9], True, False),
    ("two_parts_left_span10", [0, 1, 2, 3, 4, 6, 7, 9, 10, 11, 12, 13], False, False),
    ("two_parts_left_span9_mops", [0, 1, 2, 4, 5, 6, 7, 8, 9, 11], False, True),
])
def test_gapped_left_part_beyond_the_pair_table(name, offsets, localization, mops):
    """A gapped left part that spans 9 or 10 bases: the find kernel works from a pair table over the first 8 bases (a superset of
    the windows in the table), the expand kernel marks the false hits.  Full scan, lists, subset + recount and the OOPS scan
    against the oracle."""
    import oracle_lib as ol
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep
    ctx, octx, oset, onm, bg, bgl, n, codes = _setup(4000, 350, 29, False, mops, localization)
    try:
        total, _, mxl, _ = ctx.seqset_info()
        nent = len(n)
        prof = np.tile(hostprep.TATA_PROFILE, (1, 2))[:, :len(offsets)].copy()
        prof[:, 8:] = np.array([[55, 20, 15, 10], [15, 50, 15, 20], [20, 20, 10, 50], [10, 10, 60, 20]], dtype=np.float64)[:, :max(0, len(offsets) - 8)]
        pwm_c, cols_c = hostprep.profile_pwm(prof)
        first = 5
        cols = (first + np.asarray(offsets)).astype(np.int32)
        pwm = np.asarray(pwm_c, dtype=np.float64).reshape(31, 5).copy()
        src = pwm[cols_c[0]:cols_c[0] + len(offsets)].copy()
        pwm[:] = 0.0
        pwm[1:, 1:] = np.log(0.25)
        for k, c in enumerate(cols):
            pwm[c] = src[k]
        enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)
        ctx.reinitialize(pwm, cols, nent, int(total[nent]))
        tab = ol.Tables(pwm, cols, nent, int(total[nent]), bgl, onm)
        assert tab.error == 0 and tab.nparts == (2 if len(offsets) > 8 else 1)
        nb, nbest, mx = ctx.scan_full(nent, enr)
        octx.scan_full(tab, cols, nent, enr, True)
        assert (nb, nbest, mx) == (octx.nbSites, octx.nbBestSites, octx.maxSites) and nbest > 500
        _site_eq("best", ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_BY_SCORE), octx.best())
        _site_eq("sites", ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS), octx.sites())
        ok, nbm, bp, new = ctx.update_pwm_with_best_subset(xb.LIST_SORTED_SITES, pwm, int(cols[0]), nent, False, 0.1)
        ook, onbm, obp, onew = octx.update_pwm_with_best_subset(0, pwm, cols, nent, 0, False, 0.1)
        assert (ok, nbm, bp) == (ook, onbm, obp) and (new[5:] == onew[5:]).all()
        st = ctx.stats()
        assert st["last_scan_hits"] >= nbest  # the find kernel's hits are a superset of the candidates
    finally:
        ctx.close()
