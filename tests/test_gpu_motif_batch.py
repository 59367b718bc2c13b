"""GPU: several contexts of one device on one resident set (xxm_share_seqset / xxm_view_create) and the motifs of
MotifContainer::iterate refined several at a time (xxm_iterate_motifs) -- SURVEY.md 8(f) row 2, "batch motifs".
A view must be a full context: the golden traces of the unmodified reference replay bit-exactly on a view, and every motif
refined concurrently ends exactly where the serial loop (MotifContainer.cpp:216-270) leaves it."""
import numpy as np
import pytest

from replay import TraceInputs, replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1_zoops", "c2_mops_rc", "c5_loc_gap", "c6_long12_mops"])
def test_golden_replay_on_a_view(golden, name):
    from conftest import golden_names
    if name not in golden_names():
        pytest.skip("no such golden trace")
    from gpu_backend import GpuBackend
    from xxmotif_b200 import refine
    t = golden(name)
    ti = TraceInputs(t)
    owner = GpuBackend(ti)
    view = None
    try:
        view = refine.make_view(owner.ctx)
        assert refine.view_count(owner.ctx) == 1
        be = GpuBackend(ti, ctx=view)
        done = replay(t, be)
        assert "opt1" in done and "oops" in done and "hist" in done
        total, avg, mxl, _ = view.seqset_info()
        assert np.array_equal(total, np.asarray(ti.total).reshape(-1)) and mxl == int(t["in.pos.max_leng"][0])
    finally:
        if view is not None:
            view.close()
        assert refine.view_count(owner.ctx) == 0
        owner.ctx.close()


def _start_motifs(nent, mxl):
    """valid start models of different widths (Motif::InitProfMotif shapes): the TATA profile cut / extended to k columns, and a
    GC-rich profile that the planted instances do not carry"""
    from xxmotif_b200 import hostprep, refine, workloads
    out = []
    profs = [workloads.profile(k) for k in (8, 6, 7, 10, 12, 5)]
    gc = np.array([[5, 5, 10, 5, 5, 10], [45, 40, 70, 45, 5, 40], [45, 50, 10, 45, 85, 40], [5, 5, 10, 5, 5, 10]], dtype=np.float64)
    profs += [gc, gc[:, ::-1].copy()]
    for prof in profs:
        pwm, cols = hostprep.profile_pwm(prof)
        c0 = int(cols[0])
        out.append(lambda pwm=pwm, cols=cols, c0=c0: refine.Motif(pwm, cols, nent, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=max(1 << 20, 2 * nent)))
    return out


def _state(m):
    return (m.cols.tolist(), m.pwm.tobytes(), m.enr.tolist(), m.pval, m.posPval, m.worstScore, m.siteSeq.tobytes(), m.sitePos.tobytes())


@pytest.mark.parametrize("mops,revcomp", [(False, False), (True, True)])
def test_iterate_motifs_equals_the_serial_loop(mops, revcomp):
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, refine
    n, codes = hostprep.synthetic_set(3000, 300, seed=11)
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    owner = xb.Context(0)
    views = []
    try:
        owner.set_options(mops=mops, bg=bg, bgLog=bgl)
        owner.load_seqset(n, codes, revcomp=revcomp)
        owner.set_background(2, rp, rc)
        total, _, mxl, _ = owner.seqset_info()
        nent = len(n)
        R = refine.Refiner(owner, totalPositions=int(total[nent]), revcomp=revcomp, localization=False)
        make = _start_motifs(nent, mxl)
        # the reference's loop: one motif after the other on one context
        serial, rounds_serial = [], []
        for mk in make:
            m = mk()
            rounds_serial.append(R.iterate_motif(m))
            serial.append(_state(m))
        assert len({s[1] for s in serial}) > 1  # the motifs do end in different places
        # the same motifs, three at a time
        views = [refine.make_view(owner) for _ in range(2)]
        assert refine.view_count(owner) == 2
        for _ in range(2):  # a second pass reuses the views' buffers
            motifs = [mk() for mk in make]
            rounds = refine.iterate_motifs([owner] + views, motifs, R)
            assert rounds == rounds_serial
            for i, m in enumerate(motifs):
                assert _state(m) == serial[i], "motif %d differs from the serial loop" % i
        # one context through the batch entry point = the serial loop
        motifs = [mk() for mk in make[:3]]
        assert refine.iterate_motifs([views[1]], motifs, R) == rounds_serial[:3]
        assert [_state(m) for m in motifs] == serial[:3]
        # the owner cannot pull the set from under its views
        with pytest.raises(xb.XxmError):
            owner.load_seqset(n, codes, revcomp=revcomp)
        with pytest.raises(xb.XxmError):
            owner.set_background(2, rp, rc)
        # ... and still works afterwards; a view that loads its own set stops being one
        views[0].load_seqset(n[:100], codes[:100 * 300], revcomp=revcomp)
        assert refine.view_count(owner) == 1
        m = make[0]()
        assert R.iterate_motif(m) == rounds_serial[0] and _state(m) == serial[0]
    finally:
        for v in views:
            v.close()
        owner.close()


def _golden_all():
    from conftest import golden_names
    return golden_names()


@pytest.mark.parametrize("name", _golden_all())
def test_refine_callers_with_variants_on_views(golden, name):
    """maximizeMotifLength / the per-motif loop with the column variants spread over the context and two views
    (xxm_maximize_motif_length_on, xxm_iterate_motif_on) against the reference's own runs of those functions (stages E, F of the
    golden traces): bit-exact, like the one-context calls in test_gpu_refine.py."""
    from gpu_backend import GpuBackend
    from test_gpu_refine import _cmp_motif
    from xxmotif_b200 import localization, refine
    t = golden(name)
    ti = TraceInputs(t)
    be = GpuBackend(ti)
    views = []
    try:
        views = [refine.make_view(be.ctx) for _ in range(2)]
        pss = int(t["m0.posSetSize"][0])
        m = refine.Motif(t["m0.pwm"], t["m0.cols"], pss, enr=t["m0.enrich"], siteSeq=t["m0.sites.seq"], sitePos=t["m0.sites.pos"],
                         pval=float(t["m0.pval"][0]), posPval=float(t["m0.posPval"][0]))
        R = refine.Refiner(be.ctx, totalPositions=int(ti.total[pss]), revcomp=ti.revcomp, localization=ti.localization,
                           maximizeMotifLength=int(t["cfg.maximizeMotifLength"][0]), minMatchPositions=int(t["cfg.minMatchPositions"][0]),
                           maxMatchPositions=int(t["cfg.maxMatchPositions"][0]), pseudo=ti.pseudo, neff_pwm=float(t["cfg.neff_pwm"][0]))
        R.optimize_motif(m)
        assert np.array_equal(m.pwm[5:], t["opt1.pwm_out"][5:])
        if ti.localization:
            cols = m.cols
            e = localization.update_enriched_region(be.ctx, int(cols[0]), int(cols[-1] - cols[0] + 1), m.enr)
            for i in range(3):
                m.m.enr[i] = e[i]
        R.maximize_motif_length_on(m, views)
        _cmp_motif("maxlen on views", m, t, "maxlen")
        rounds = R.iterate_motif(m, views=views)
        assert rounds >= 1
        _cmp_motif("converged on views", m, t, "converged")
    finally:
        for v in views:
            v.close()
        be.ctx.close()
