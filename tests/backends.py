"""Backends for tests/replay.py: the C restatement (OracleBackend).  The CUDA backend lives in
tests/gpu_backend.py so that importing this module never needs a GPU."""
import numpy as np

import oracle_lib as ol


def tables_dump_dense(L, maxk, nparts, pvals, probs, classes):
    d = {"L": L, "maxk": maxk, "nparts": nparts, "ids": [], "vals": [], "cscore": [], "ccount": [], "cids": []}
    for part in range(nparts):
        pv = pvals[part]
        ids = np.flatnonzero(pv != 0).astype(np.int64)
        d["ids"].append(ids)
        d["vals"].append(pv[ids])
        sc, cnt, cid = classes[part]
        d["cscore"].append(sc)
        d["ccount"].append(cnt)
        d["cids"].append(cid)
    if nparts == 2:
        ids = np.flatnonzero(probs != 0).astype(np.int64)
        d["prob_ids"] = ids
        d["prob_vals"] = probs[ids]
    return d


class OracleBackend:
    """Drives oracle/liboracle.so with the inputs of one trace."""

    def __init__(self, ti, from_scratch=True):
        self.ti = ti
        n_pre, S_pre = ti.pos_pre()
        self.seqset = ol.SeqSet(n_pre, S_pre, revcomp=bool(ti.revcomp))
        t = ti.t
        if from_scratch:
            if ti.hasNeg:
                nn, ns = ti.neg_pre()
                self.negset = ol.SeqSet(nn, ns, revcomp=bool(ti.revcomp), recount=True)
                bgset = self.negset
            else:
                bgset = self.seqset
            self.bg, self.bgLog = bgset.background()
            self.nm = ol.NullModel(bgset, alpha=float(t["cfg.pseudocountsFactor"][0]), beta=float(t["cfg.countsOffset"][0]), order=ti.order)
        else:
            self.bg, self.bgLog = ti.bg, ti.bgLog
            self.nm = ol.NullModel(order=ti.order, rprobs=t["nm.rprobs"], rconds=t["nm.rconds"])
        self.ctx = ol.Ctx(self.seqset, ti.revcomp, ti.mops, ti.oops, ti.localization, self.bg, self.bgLog,
                          maxMotifsPerSequence=int(t["cfg.maxMotifsPerSequence"][0]), maxSeqCount=int(t["cfg.maxSeqCount"][0]),
                          plusFrac=float(t["cfg.plusFrac"][0]), overrepCorrection=float(t["cfg.overrepCorrection"][0]))

    def tables(self, pwm, cols, posSetSize, totalPositions):
        h = ol.Tables(pwm, cols, posSetSize, totalPositions, self.bgLog, self.nm)
        assert h.error == 0
        return h

    def tables_dump(self, h):
        np_ = h.nparts
        return tables_dump_dense(h.ncols, h.maxk, np_, [h.pvalues(p) for p in range(np_)], h.probs() if np_ == 2 else None,
                                 [h.classes(p) for p in range(np_)])

    def window_pvals(self, h, cols, posSetSize, enr, seq, pos, mops, firstMotifColumn=None):
        length = int(cols[-1] - cols[0] + 1)
        if self.ti.localization:
            fmc = int(cols[0]) if firstMotifColumn is None else firstMotifColumn
            self.ctx.init_positional_probs(length, fmc, enr, int(cols[0]) - fmc)
        f = np.zeros(len(seq))
        s = np.zeros(len(seq))
        for k in range(len(seq)):
            f[k], s[k] = self.ctx.calculate_pval(h, int(seq[k]), int(pos[k]), length, enr, posSetSize, mops)
        return f, s

    def scan_full(self, h, cols, posSetSize, enr):
        self.ctx.scan_full(h, cols, posSetSize, enr, True)
        return {"best": self.ctx.best(), "sites": self.ctx.sites(), "nbSites": self.ctx.nbSites, "nbBestSites": self.ctx.nbBestSites,
                "maxSites": self.ctx.maxSites}

    def update_pwm(self, which, pwm, cols, posSetSize, plus, pseudo):
        ok, nb, bp, out = self.ctx.update_pwm_with_best_subset(which, pwm, cols, posSetSize, 0, plus, pseudo)
        srt = self.ctx.elong() if which else self.ctx.sites()
        return {"ok": ok, "bestMotifNb": nb, "bestPval": bp, "pwm": out, "sorted": srt}

    def rescore(self, h, firstMotifColumn, tmpCols, posSetSize, enr):
        self.ctx.rescore(h, firstMotifColumn, tmpCols, posSetSize, enr, False)
        return {"elong": self.ctx.elong(), "nbSites": self.ctx.nbSites, "maxSites": self.ctx.maxSites}

    def rescore_startposlist(self, h, cols, posSetSize, enr, seq, pos):
        self.ctx.rescore_startposlist(h, cols, posSetSize, enr, seq, pos)
        return {"sites": self.ctx.sites(), "nbSites": self.ctx.nbSites, "maxSites": self.ctx.maxSites}

    def startposlist_best_subset(self, cols, posSetSize):
        n, nb, bp, s, p = self.ctx.startposlist_best_subset(cols, posSetSize, 0)
        return {"n": n, "bestMotifNb": nb, "bestPval": bp, "seq": s, "pos": p}

    def scan_oops(self, h, cols, posSetSize, enr):
        self.ctx.scan_oops(h, cols, posSetSize, enr)
        return {"sites": self.ctx.sites(self.seqset.nent), "nbSites": self.ctx.nbSites, "maxSites": self.ctx.maxSites}

    def binding_sites(self, thr):
        n = self.ctx.set_binding_sites(thr)
        return self.ctx.sites(n)

    def logodds_hist(self, pwm, cols, posSetSize):
        return self.ctx.logodds_hist(pwm, cols, posSetSize)
