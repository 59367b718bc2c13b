"""Replay a golden boundary trace (tests/golden/*.trace.gz, dumped from the unmodified reference by
oracle/ref_harness.cpp) against a backend and compare every stage bit-exactly.

The same replay drives the C restatement (tests/oracle_lib.py) and the CUDA path (xxmotif_b200), so the GPU
parity tests read exactly like the oracle's own pin.

A backend provides:
    tables(pwm155, cols, posSetSize, totalPositions)                 -> handle
    tables_dump(handle)                                              -> dict  (see _cmp_tables)
    window_pvals(handle, cols, posSetSize, enr, seq[], pos[], mops, firstMotifColumn=None)
                                                                     -> (finalPval[], pValScore[])  float64
    scan_full(handle, cols, posSetSize, enr)                         -> dict(best, sites, nbSites, nbBestSites, maxSites)
    update_pwm(which, pwm155, cols, posSetSize, plus, pseudo)        -> dict(ok, bestMotifNb, bestPval, pwm, sorted)
    rescore(handle, firstMotifColumn, tmpCols, posSetSize, enr)      -> dict(elong, nbSites, maxSites)
    rescore_startposlist(handle, cols, posSetSize, enr, seq[], pos[])-> dict(sites, nbSites, maxSites)
    startposlist_best_subset(cols, posSetSize)                       -> dict(n, bestMotifNb, bestPval, seq, pos)
    scan_oops(handle, cols, posSetSize, enr)                         -> dict(sites, nbSites, maxSites)
    binding_sites(thr)                                               -> structured site array sorted by start
    logodds_hist(pwm155, cols, posSetSize)                           -> (fwd, bwd, max1, max2)
site arrays are numpy structured arrays with fields score(f4), startPos(i4), sequence(i4).
"""
import numpy as np


def _eq(name, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s != %s" % (name, a.shape, b.shape)
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not same.all():
        idx = np.flatnonzero(~same.reshape(-1))[:5]
        raise AssertionError("%s: %d mismatches, first at %s: got %s want %s" % (
            name, int((~same).sum()), idx, a.reshape(-1)[idx], b.reshape(-1)[idx]))


def _cmp_sites(name, got, t, prefix, n=None, mask_unset=False):
    seq, pos, sc = t[prefix + ".seq"], t[prefix + ".pos"], t[prefix + ".score"]
    if n is not None:
        seq, pos, sc = seq[:n], pos[:n], sc[:n]
        got = got[:n]
    _eq(name + ".score", got["score"], sc)
    if mask_unset:  # ZOOPS slots with score 1 hold stale (seq,pos) from earlier calls in the reference
        m = sc != 1
        _eq(name + ".seq", got["sequence"][m], seq[m])
        _eq(name + ".pos", got["startPos"][m], pos[m])
    else:
        _eq(name + ".seq", got["sequence"], seq)
        _eq(name + ".pos", got["startPos"], pos)


def _cmp_tables(name, d, t, p):
    _eq(name + ".L", d["L"], int(t[p + ".L"][0]))
    _eq(name + ".maxk", d["maxk"], int(t[p + ".maxk"][0]))
    nparts = int(t[p + ".hashNumber"][0])
    _eq(name + ".nparts", d["nparts"], nparts)
    for part in range(nparts):
        q = "%s.p%d" % (p, part)
        _eq(name + ".p%d.ids" % part, d["ids"][part], t[q + ".ids"])
        _eq(name + ".p%d.vals" % part, d["vals"][part], t[q + ".vals"])
        q = "%s.c%d" % (p, part)
        _eq(name + ".c%d.score" % part, d["cscore"][part], t[q + ".score"])
        _eq(name + ".c%d.count" % part, d["ccount"][part], t[q + ".count"])
        _eq(name + ".c%d.ids" % part, d["cids"][part], t[q + ".ids"])
    if nparts == 2:
        _eq(name + ".prob.ids", d["prob_ids"], t[p + ".prob.ids"])
        _eq(name + ".prob.vals", d["prob_vals"], t[p + ".prob.vals"])


class TraceInputs:
    """Decoded inputs of a trace."""

    def __init__(self, t):
        self.t = t
        g = lambda k: int(t[k][0])
        self.revcomp, self.mops, self.oops = g("cfg.revcomp"), g("cfg.mops"), g("cfg.oops")
        self.localization, self.order = g("cfg.localization"), g("nm.order")
        self.pseudo = float(t["cfg.pseudo"][0])
        self.hasNeg = g("in.hasNeg")
        self.n_post = t["in.pos.n"]
        self.S_post = t["in.pos.S"]
        self.total = t["in.pos.total"]
        self.avgLength = t["in.pos.avgLength"]
        self.bg = t["bg.negBg"] if self.hasNeg else t["bg.posBg"]
        self.bgLog = t["bg.negBg_log"] if self.hasNeg else t["bg.posBg_log"]

    @staticmethod
    def _pre(n_post, S_post, revcomp):
        if not revcomp:
            return n_post.copy(), S_post.copy()
        n_pre = (n_post - 1) // 2
        out = []
        off = 0
        for a, b in zip(n_pre, n_post):
            out.append(S_post[off:off + a])
            off += b
        return n_pre.astype(np.int32), (np.concatenate(out) if out else np.zeros(0, np.uint8))

    def pos_pre(self):
        return self._pre(self.n_post, self.S_post, self.revcomp)

    def neg_pre(self):
        return self._pre(self.t["in.neg.n"], self.t["in.neg.S"], self.revcomp)


def replay(t, be, check_tables=True, check_hist=True):
    """Run every stage of trace `t` through backend `be`; raises AssertionError on the first mismatch."""
    ti = TraceInputs(t)
    pseudo = ti.pseudo
    total = ti.total
    done = []

    def totpos(posSetSize):
        return int(total[posSetSize])

    # ---- A. optimizeMotif body ------------------------------------------------------------------
    enr = None
    for it in range(2):
        p = "opt%d" % it
        pwm, cols = t[p + ".in.pwm"], t[p + ".in.cols"]
        pss = int(t[p + ".in.posSetSize"][0])
        e = t[p + ".in.enrich"]
        enr = (int(e[1]), int(e[2]))
        h = be.tables(pwm, cols, pss, totpos(pss))
        if check_tables:
            _cmp_tables(p + ".tab", be.tables_dump(h), t, p + ".tab")
        for tag, mops in (("lk_mops", 1), ("lk_mode", ti.mops)):
            f, s = be.window_pvals(h, cols, pss, enr, t[p + "." + tag + ".seq"], t[p + "." + tag + ".pos"], mops)
            _eq(p + "." + tag + ".pValScore", s, t[p + "." + tag + ".pValScore"])
            _eq(p + "." + tag + ".finalPval", f, t[p + "." + tag + ".finalPval"])
        r = be.scan_full(h, cols, pss, enr)
        _eq(p + ".nbSites", r["nbSites"], int(t[p + ".nbSites"][0]))
        _eq(p + ".nbBestSites", r["nbBestSites"], int(t[p + ".nbBestSites"][0]))
        _eq(p + ".maxSites", r["maxSites"], int(t[p + ".maxSites"][0]))
        _cmp_sites(p + ".best", r["best"], t, p + ".best")
        _cmp_sites(p + ".sites", r["sites"], t, p + ".sites")
        u = be.update_pwm(0, pwm, cols, pss, False, pseudo)
        if int(t[p + ".nbSites"][0]) > 0 and t[p + ".sorted.score"][0] != 1:
            _eq(p + ".bestMotifNb", u["bestMotifNb"], int(t[p + ".bestMotifNb"][0]))
            _eq(p + ".bestPval", u["bestPval"], float(t[p + ".bestPval"][0]))
            _cmp_sites(p + ".sorted", u["sorted"], t, p + ".sorted")
            _eq(p + ".pwm_out", u["pwm"][5:], t[p + ".pwm_out"][5:])
        done.append(p)
    motif_cols = t["opt1.in.cols"]
    pss = int(t["opt1.in.posSetSize"][0])
    if ti.localization:
        e = t["optA.enriched.enrich"]
        enr = (int(e[1]), int(e[2]))

    # ---- D. candidate re-scoring for every column variant ---------------------------------------
    for v in range(int(t["el.count"][0])):
        for it in range(3):
            p = "el%d.%d" % (v, it)
            if p + ".cols" not in t:
                break
            tmp = t[p + ".cols"]
            pwm = t[p + ".pwm_in"]
            h = be.tables(pwm, tmp, pss, totpos(pss))
            if it == 0:
                if check_tables:
                    _cmp_tables(p + ".tab", be.tables_dump(h), t, p + ".tab")
                f, s = be.window_pvals(h, tmp, pss, enr, t[p + ".lk_mode.seq"], t[p + ".lk_mode.pos"], ti.mops,
                                       firstMotifColumn=int(motif_cols[0]))
                _eq(p + ".lk_mode.pValScore", s, t[p + ".lk_mode.pValScore"])
                _eq(p + ".lk_mode.finalPval", f, t[p + ".lk_mode.finalPval"])
            r = be.rescore(h, int(motif_cols[0]), tmp, pss, enr)
            nb = int(t[p + ".nbSites"][0])
            _eq(p + ".nbSites", r["nbSites"], nb)
            _eq(p + ".maxSites", r["maxSites"], int(t[p + ".maxSites"][0]))
            _cmp_sites(p + ".elong", r["elong"], t, p + ".elong", n=nb, mask_unset=not ti.mops)
            u = be.update_pwm(1, pwm, tmp, pss, it == 0, pseudo)
            if nb > 0 and t[p + ".sorted.score"][0] != 1:
                _eq(p + ".motifNb", u["bestMotifNb"], int(t[p + ".motifNb"][0]))
                _eq(p + ".overrepPval", u["bestPval"], float(t[p + ".overrepPval"][0]))
                _cmp_sites(p + ".sorted", u["sorted"], t, p + ".sorted", n=nb, mask_unset=not ti.mops)
                _eq(p + ".pwm_out", u["pwm"][5:], t[p + ".pwm_out"][5:])
            done.append(p)

    # ---- J. re-score the motif's own site list (Motif::updatePval body) --------------------------
    pwm, cols = t["converged.pwm"], t["converged.cols"]
    pss = int(t["converged.posSetSize"][0])
    e = t["converged.enrich"]
    enr = (int(e[1]), int(e[2]))
    h = be.tables(pwm, cols, pss, totpos(pss))
    if check_tables:
        _cmp_tables("upd.tab", be.tables_dump(h), t, "upd.tab")
    r = be.rescore_startposlist(h, cols, pss, enr, t["converged.sites.seq"], t["converged.sites.pos"])
    nb = int(t["upd.nbSites"][0])
    _eq("upd.nbSites", r["nbSites"], nb)
    _eq("upd.maxSites", r["maxSites"], int(t["upd.maxSites"][0]))
    _cmp_sites("upd.sites", r["sites"], t, "upd.sites", n=nb, mask_unset=not ti.mops)
    if nb > 0:
        b = be.startposlist_best_subset(cols, pss)
        if b["n"] >= 0:
            _eq("upd.bestMotifNb", b["bestMotifNb"], int(t["upd.bestMotifNb"][0]))
            _eq("upd.bestPval", b["bestPval"], float(t["upd.bestPval"][0]))
            _eq("upd.after.sites.seq", b["seq"], t["upd.after.sites.seq"])
            _eq("upd.after.sites.pos", b["pos"], t["upd.after.sites.pos"])
    done.append("upd")

    # ---- G. final OOPS scan + binding sites --------------------------------------------------------
    h = be.tables(pwm, cols, pss, totpos(pss))
    r = be.scan_oops(h, cols, pss, enr)
    _eq("oops.nbSites", r["nbSites"], int(t["oops.nbSites"][0]))
    _eq("oops.maxSites", r["maxSites"], int(t["oops.maxSites"][0]))
    _cmp_sites("oops.sites", r["sites"], t, "oops.sites")
    bs = be.binding_sites(1.0)
    _eq("oops.bind.seq", bs["sequence"], t["oops.bind.seq"])
    _eq("oops.bind.pos", bs["startPos"] - int(cols[0]) + 1, t["oops.bind.pos"])
    _eq("oops.bind.score", bs["score"].astype(np.float64), t["oops.bind.score"])
    done.append("oops")

    # ---- H. per-site p-values ----------------------------------------------------------------------
    sseq, spos = t["upd.after.sites.seq"], t["upd.after.sites.pos"]
    if len(sseq):
        f, _ = be.window_pvals(h, cols, pss, enr, sseq, spos + int(cols[0]) - 1, ti.mops)
        _eq("persite.prob", f, t["persite.prob"])
    done.append("persite")

    # ---- I. forward / reverse-complement log-odds histogram ----------------------------------------
    if check_hist:
        fw, bw, m1, m2 = be.logodds_hist(pwm, cols, pss)
        _eq("hist.maxScore", m1, float(t["hist.maxScore"][0]))
        _eq("hist.maxScore2", m2, float(t["hist.maxScore2"][0]))
        _eq("hist.fwd", fw[:len(t["hist.fwd"])], t["hist.fwd"])
        _eq("hist.bwd", bw[:len(t["hist.bwd"])], t["hist.bwd"])
        done.append("hist")
    return done
