"""tests/replay.py backend that drives the CUDA path through the C ABI (xxmotif_b200.capi)."""
import numpy as np

import xxmotif_b200 as xb
from backends import tables_dump_dense


class GpuBackend:
    def __init__(self, ti, device=0, ctx=None):
        self.ti = ti
        t = ti.t
        if ctx is not None:  # a context that already holds the trace's set, options and background (a view, refine.make_view)
            self.ctx = ctx
            self.nent = ctx.nent
            self.max_leng = int(t["in.pos.max_leng"][0])
            return
        self.ctx = xb.Context(device)
        self.ctx.set_options(mops=ti.mops, oops=ti.oops, localization=ti.localization, bg=ti.bg, bgLog=ti.bgLog,
                             maxMotifsPerSequence=int(t["cfg.maxMotifsPerSequence"][0]), maxSeqCount=int(t["cfg.maxSeqCount"][0]),
                             plusFrac=float(t["cfg.plusFrac"][0]), overrepCorrection=float(t["cfg.overrepCorrection"][0]))
        n_pre, S_pre = ti.pos_pre()
        self.ctx.load_seqset(n_pre, S_pre, revcomp=bool(ti.revcomp))
        self.nent = len(n_pre)
        self.max_leng = int(t["in.pos.max_leng"][0])
        self.ctx.set_background(ti.order, self.rprobs(), self.rconds())

    # NullModel tables are an INPUT of the path (NullModel::init is row f-3, "next"): from the trace, or for the
    # order-8 case (too large for a fixture) from the C restatement, which test_oracle pins by hash.
    def _nm(self):
        if not hasattr(self, "_nm_cache"):
            t = self.ti.t
            if "nm.rprobs" in t:
                self._nm_cache = (t["nm.rprobs"], t["nm.rconds"])
            else:
                import oracle_lib as ol
                nn, ns = self.ti.neg_pre() if self.ti.hasNeg else self.ti.pos_pre()
                bgset = ol.SeqSet(nn, ns, revcomp=bool(self.ti.revcomp), recount=bool(self.ti.hasNeg))
                nm = ol.NullModel(bgset, alpha=float(t["cfg.pseudocountsFactor"][0]), beta=float(t["cfg.countsOffset"][0]), order=self.ti.order)
                self._nm_cache = (nm.rprobs, nm.rconds)
        return self._nm_cache

    def rprobs(self):
        return self._nm()[0]

    def rconds(self):
        return self._nm()[1]

    def tables(self, pwm, cols, posSetSize, totalPositions):
        self.ctx.reinitialize(pwm, cols, posSetSize, totalPositions)
        return None

    def tables_dump(self, h):
        d = self.ctx.tables()
        return tables_dump_dense(d["L"], d["maxk"], d["nparts"], d["pvalues"], d["probs"], d["classes"])

    def window_pvals(self, h, cols, posSetSize, enr, seq, pos, mops, firstMotifColumn=None):
        fmc = int(cols[0]) if firstMotifColumn is None else firstMotifColumn
        return self.ctx.window_pvals(seq, pos, posSetSize, fmc, enr, mops)

    def scan_full(self, h, cols, posSetSize, enr):
        nb, nbest, mx = self.ctx.scan_full(posSetSize, enr)
        return {"best": self.ctx.get_sites(xb.LIST_BEST_STARTPOS, xb.ORDER_BY_SCORE), "sites": self.ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS),
                "nbSites": nb, "nbBestSites": nbest, "maxSites": mx}

    def update_pwm(self, which, pwm, cols, posSetSize, plus, pseudo):
        ok, nb, bp, out = self.ctx.update_pwm_with_best_subset(which, pwm, int(cols[0]), posSetSize, plus, pseudo)
        return {"ok": ok, "bestMotifNb": nb, "bestPval": bp, "pwm": out, "sorted": self.ctx.get_sites(which, xb.ORDER_BY_SCORE)}

    def rescore(self, h, firstMotifColumn, tmpCols, posSetSize, enr):
        nb, mx = self.ctx.rescore(firstMotifColumn, posSetSize, enr)
        return {"elong": self.ctx.get_sites(xb.LIST_SORTED_SITES_ELONG, xb.ORDER_AS_IS), "nbSites": nb, "maxSites": mx}

    def rescore_startposlist(self, h, cols, posSetSize, enr, seq, pos):
        nb, mx = self.ctx.rescore_startposlist(posSetSize, enr, seq, pos)
        return {"sites": self.ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS), "nbSites": nb, "maxSites": mx}

    def startposlist_best_subset(self, cols, posSetSize):
        n, nb, bp, s, p = self.ctx.startposlist_best_subset(int(cols[0]), posSetSize, cap=1 << 16)
        return {"n": n, "bestMotifNb": nb, "bestPval": bp, "seq": s, "pos": p}

    def scan_oops(self, h, cols, posSetSize, enr):
        nb, mx = self.ctx.scan_oops(posSetSize, enr)
        return {"sites": self.ctx.get_sites(xb.LIST_SORTED_SITES, xb.ORDER_AS_IS), "nbSites": nb, "maxSites": mx}

    def binding_sites(self, thr):
        return self.ctx.set_binding_sites(thr, cap=max(1, self.nent))

    def logodds_hist(self, pwm, cols, posSetSize):
        return self.ctx.logodds_hist(pwm, cols, posSetSize, max(0, self.max_leng - len(cols) + 1))
