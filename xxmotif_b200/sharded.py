"""One refinement step over a sequence set sharded across ranks (one process per GPU).

Sequences are independent units: every rank scans and de-overlaps its own shard with no data-path collective.  The
path has one real exchange per step -- the optimal subset is chosen over ALL sites and the PWM is recounted from
them -- so only the run-length histogram of site scores (all-gather) and the 31x5 count matrix (all-reduce) cross
NVLink.  With `dist=None` the same split-phase calls run on one GPU.
"""
import ctypes as C

import numpy as np

from . import capi

EXTRA_EXPORTS = ["xxm_score_runs", "xxm_count_sites", "xxm_host_choose_subset", "xxm_host_build_pwm"]


def _lib():
    L = capi.load_library()
    if not getattr(L, "_sharded_ready", False):
        vp, i32, i64, dbl, flt = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float
        P = C.POINTER
        L.xxm_score_runs.restype = i32
        L.xxm_score_runs.argtypes = [vp, i32, vp, vp, i32, P(i32)]
        L.xxm_count_sites.restype = i32
        L.xxm_count_sites.argtypes = [vp, i32, flt, i32, vp, P(i64)]
        L.xxm_host_choose_subset.restype = i32
        L.xxm_host_choose_subset.argtypes = [vp, vp, i32, i64, i64, i64, i32, dbl, P(i64), P(dbl), P(flt)]
        L.xxm_host_build_pwm.restype = i32
        L.xxm_host_build_pwm.argtypes = [vp, dbl, vp, vp]
        L._sharded_ready = True
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def score_runs(ctx, which, cap=1 << 16):
    L = _lib()
    while True:
        sc = np.zeros(cap, dtype=np.float32)
        cn = np.zeros(cap, dtype=np.int64)
        n = C.c_int32(0)
        rc = L.xxm_score_runs(ctx.h, which, _p(sc), _p(cn), cap, C.byref(n))
        if rc and n.value > cap:
            cap = n.value
            continue
        ctx._ck(rc)
        return sc[:n.value].copy(), cn[:n.value].copy()


def count_sites(ctx, which, thr, firstColumn):
    L = _lib()
    counts = np.zeros(155, dtype=np.float64)
    ns = C.c_int64(0)
    ctx._ck(L.xxm_count_sites(ctx.h, which, float(thr), firstColumn, _p(counts), C.byref(ns)))
    return counts


def merge_runs(score_lists, count_lists):
    """Merge per-rank (ascending score, count) runs into global runs."""
    if not score_lists:
        return np.zeros(0, np.float32), np.zeros(0, np.int64)
    sc = np.concatenate(score_lists).astype(np.float32)
    cn = np.concatenate(count_lists).astype(np.int64)
    if sc.size == 0:
        return sc, cn
    order = np.argsort(sc, kind="stable")
    sc, cn = sc[order], cn[order]
    first = np.concatenate(([True], sc[1:] != sc[:-1]))
    idx = np.flatnonzero(first)
    return sc[idx], np.add.reduceat(cn, idx)


def choose_subset(scores, counts, nbSites, limit, maxSites, oneOccurrence=False, plusFrac=0.0):
    L = _lib()
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    nb, bp, ls = C.c_int64(0), C.c_double(1.0), C.c_float(0.0)
    rc = L.xxm_host_choose_subset(_p(scores), _p(counts), len(scores), int(nbSites), int(limit), int(maxSites), int(oneOccurrence), plusFrac,
                                  C.byref(nb), C.byref(bp), C.byref(ls))
    if rc:
        raise capi.XxmError("xxm_host_choose_subset failed with %d" % rc)
    return nb.value, bp.value, ls.value


def build_pwm(counts, pseudoTimesK, bg):
    L = _lib()
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    bg = np.ascontiguousarray(bg, dtype=np.float64)
    pwm = np.zeros(155, dtype=np.float64)
    L.xxm_host_build_pwm(_p(counts), float(pseudoTimesK), _p(bg), _p(pwm))
    return pwm


class Exchange:
    """The two small collectives of a step over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

    Both are latency bound (a few KB), so each is ONE collective on preallocated buffers: the runs of a rank travel as one
    fixed-size int64 row [n, scalars..., scores as int32 bit patterns..., counts...] through all_gather_into_tensor, the
    31x5 counts through one all_reduce; host <-> device staging goes through pinned memory."""

    CAP_RUNS = 4096
    NSCAL = 3

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.device = device
        self.world = dist.get_world_size() if dist else 1
        self._ready = False

    def _setup(self):
        import torch
        cuda = self.device is not None and self.device.type == "cuda"
        W = 1 + self.NSCAL + 2 * self.CAP_RUNS
        self._W = W
        self._h_in = torch.zeros(W, dtype=torch.int64, pin_memory=cuda)
        self._h_out = torch.zeros(self.world * W, dtype=torch.int64, pin_memory=cuda)
        self._d_in = torch.zeros(W, dtype=torch.int64, device=self.device)
        self._d_out = torch.zeros(self.world * W, dtype=torch.int64, device=self.device)
        self._h_cnt = torch.zeros(155, dtype=torch.float64, pin_memory=cuda)
        self._d_cnt = torch.zeros(155, dtype=torch.float64, device=self.device)
        self._cuda = cuda
        self._ready = True

    def gather_runs(self, sc, cn, scalars):
        """all-gather of (scores, counts) runs + sum of a few int64 scalars."""
        if not self.dist:
            return [sc], [cn], list(scalars)
        import torch
        if not self._ready:
            self._setup()
        k = len(sc)
        assert len(scalars) == self.NSCAL
        if k > self.CAP_RUNS:
            return self._gather_runs_any_size(sc, cn, scalars)
        W, C0 = self._W, 1 + self.NSCAL
        h = self._h_in.numpy()
        h[0] = k
        h[1:C0] = scalars
        h[C0:C0 + k] = sc.view(np.int32)
        h[C0 + self.CAP_RUNS:C0 + self.CAP_RUNS + k] = cn
        self._d_in.copy_(self._h_in, non_blocking=True)
        self.dist.all_gather_into_tensor(self._d_out, self._d_in)
        self._h_out.copy_(self._d_out, non_blocking=True)
        if self._cuda:
            torch.cuda.current_stream().synchronize()
        o = self._h_out.numpy().reshape(self.world, W)
        if (o[:, 0] < 0).any():  # some rank had more runs than a row holds: everybody takes the general path
            return self._gather_runs_any_size(sc, cn, scalars)
        scs = [o[r, C0:C0 + int(o[r, 0])].astype(np.int32).view(np.float32) for r in range(self.world)]
        cns = [o[r, C0 + self.CAP_RUNS:C0 + self.CAP_RUNS + int(o[r, 0])].copy() for r in range(self.world)]
        tot = [int(o[:, 1 + j].sum()) for j in range(self.NSCAL)]
        return scs, cns, tot

    def _gather_runs_any_size(self, sc, cn, scalars):
        import torch
        dist = self.dist
        if len(sc) > self.CAP_RUNS:  # tell the ranks that took the fixed-size path
            self._h_in.numpy()[0] = -1
            self._d_in.copy_(self._h_in)
            dist.all_gather_into_tensor(self._d_out, self._d_in)
        lens = torch.tensor([len(sc)] + [int(x) for x in scalars], dtype=torch.int64, device=self.device)
        all_lens = [torch.zeros_like(lens) for _ in range(self.world)]
        dist.all_gather(all_lens, lens)
        mx = max(1, max(int(t[0]) for t in all_lens))
        buf = torch.zeros(2 * mx, dtype=torch.int64, device=self.device)
        buf[:len(sc)] = torch.from_numpy(sc.view(np.int32).astype(np.int64)).to(self.device)
        buf[mx:mx + len(cn)] = torch.from_numpy(cn).to(self.device)
        bufs = [torch.zeros_like(buf) for _ in range(self.world)]
        dist.all_gather(bufs, buf)
        scs, cns = [], []
        for t, b in zip(all_lens, bufs):
            k = int(t[0])
            b = b.cpu().numpy()
            scs.append(b[:k].astype(np.int32).view(np.float32))
            cns.append(b[mx:mx + k].copy())
        tot = [sum(int(t[1 + j]) for t in all_lens) for j in range(len(scalars))]
        return scs, cns, tot

    def sum_scalars(self, vals):
        """sum of a few doubles over the ranks"""
        if not self.dist:
            return list(vals)
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]

    def max_scalar(self, v):
        if not self.dist:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_counts(self, counts):
        if not self.dist:
            return counts
        import torch
        if not self._ready:
            self._setup()
        self._h_cnt.numpy()[:] = counts
        self._d_cnt.copy_(self._h_cnt, non_blocking=True)
        self.dist.all_reduce(self._d_cnt, op=self.dist.ReduceOp.SUM)
        self._h_cnt.copy_(self._d_cnt, non_blocking=True)
        if self._cuda:
            torch.cuda.current_stream().synchronize()
        return self._h_cnt.numpy().copy()


class ShardedRefinement:
    """optimizeMotif iterations (Iterate_Motif.cpp:26-32) over a sharded set; `ctx` is the rank's xxmotif_b200.Context."""

    def __init__(self, ctx, dist, nent_local, total_local, pseudo=0.1, revcomp=False, oneOccurrence=False, maxSeqCount=1000, plusFrac=0.0,
                 bg=None, native=True, comm_ready=False):
        """native=True: the whole step is one C-ABI call (xxm_optimize_step) and the two small reductions run over NCCL on
        the library's stream, device to device.  native=False: the split-phase entry points with the exchange in Python
        (any torch.distributed backend; this is also what the CPU/gloo test drives)."""
        self.ctx = ctx
        device = None
        self.native = native
        if dist is not None:
            import torch
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            if dist.get_backend() != "nccl":
                self.native = False
        self.ex = Exchange(dist, device)
        self.nent_local, self.total_local = nent_local, total_local
        self.pseudo, self.revcomp = pseudo, revcomp
        self.oneOccurrence, self.maxSeqCount, self.plusFrac = oneOccurrence, maxSeqCount, plusFrac
        self.bg = bg
        self.last_d2h_bytes = 0
        self._global_set = False
        if self.native and dist is not None and dist.get_world_size() > 1 and not comm_ready:  # comm_ready: the context already has its communicator
            import torch
            idt = torch.zeros(128, dtype=torch.uint8, device=device)
            if dist.get_rank() == 0:
                idt.copy_(torch.from_numpy(capi.comm_unique_id()))
            dist.broadcast(idt, src=0)
            ctx.comm_init(dist.get_rank(), dist.get_world_size(), idt.cpu().numpy())

    def optimize_step(self, pwm, cols, enr):
        ctx = self.ctx
        if self.native:
            new, info = ctx.optimize_step(pwm, cols, self.nent_local, self.total_local, enr, self.pseudo)
            self.last_d2h_bytes = 8 * (ctx.nval_hint() + 3) + 155 * 8 + 64
            return new, info
        if self.ex.world > 1 and not self._global_set:
            # the tables, maxSites and the ZOOPS exponent follow the WHOLE set (branch_and_bound-inl.h:186-187, :417): sum the
            # counts over the ranks once and hand them to the context (xxm_set_global_counts)
            _, _, mxl, _ = ctx.seqset_info()
            g = self.ex.sum_scalars([float(self.nent_local), float(self.total_local), ctx.seqset_logsum()])
            gmx = self.ex.max_scalar(float(mxl))
            ctx.set_global_counts(int(round(g[0])), int(round(g[1])), g[2], int(round(gmx)))
            self._global_set = True
        ctx.reinitialize(pwm, cols, self.nent_local, self.total_local)
        nbSites, nbBest, _ = ctx.scan_full(self.nent_local, enr)
        sc, cn = score_runs(ctx, capi.LIST_SORTED_SITES)
        scs, cns, tot = self.ex.gather_runs(sc, cn, [nbSites, self.nent_local, self.total_local])
        gsc, gcn = merge_runs(scs, cns)
        nb_g, nent_g, total_g = tot
        span = int(cols[-1] - cols[0] + 1)
        maxSites = total_g - (2 if self.revcomp else 1) * nent_g * (span - 1)  # StartPosUpdater-inl.h:135-138
        info = {"nbSites": nbSites, "nbBestSites": nbBest, "nbSites_global": nb_g, "maxSites": maxSites, "bestMotifNb": 0, "bestPval": 1.0}
        self.last_d2h_bytes = 12 * len(sc) + 155 * 8 + 64
        if nb_g == 0 or len(gsc) == 0:
            return np.asarray(pwm, dtype=np.float64).reshape(155).copy(), info
        K, bestPval, thr = choose_subset(gsc, gcn, nb_g, nent_g * self.maxSeqCount, maxSites, self.oneOccurrence, self.plusFrac)
        counts = self.ex.sum_counts(count_sites(ctx, capi.LIST_SORTED_SITES, thr, int(cols[0])))
        bg = self.bg if self.bg is not None else np.array(list(ctx_bg(ctx)))
        new = build_pwm(counts, self.pseudo * K, bg)
        info["bestMotifNb"], info["bestPval"] = K, bestPval
        return new, info


def ctx_bg(ctx):
    raise capi.XxmError("ShardedRefinement needs bg= (the mono-nucleotide background given to set_options)")
