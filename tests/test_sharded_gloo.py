"""The multi-rank exchange of a refinement step (xxmotif_b200/sharded.py) with world_size 2 over gloo on CPU: gathering
and merging score runs, the single subset choice and the count all-reduce must reproduce the single-process result.
The per-rank device calls are replaced by the oracle (test infrastructure) so that no GPU is needed."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_shard(n, codes, lo, hi, length, pwm, cols, bg, bgl, rp, rc):
    """(runs, nbSites, kept sites) of the sequences [lo, hi) computed by the oracle."""
    import oracle_lib as ol
    oset = ol.SeqSet(n[lo:hi], codes[lo * length:hi * length])
    onm = ol.NullModel(order=2, rprobs=rp, rconds=rc)
    octx = ol.Ctx(oset, False, True, False, False, bg, bgl)
    tab = ol.Tables(pwm, cols, hi - lo, (hi - lo) * length, bgl, onm)
    octx.scan_full(tab, cols, hi - lo, (1 - int(cols[0]) + 1, length - int(cols[0]) + 1), True)
    sites = octx.sites()
    sc, cn = np.unique(sites["score"], return_counts=True)
    return sc.astype(np.float32), cn.astype(np.int64), sites, oset


def _counts(sites, oset, thr, firstCol):
    cnt = np.zeros((31, 5))
    S = oset.flat().reshape(oset.nent, -1)
    L = S.shape[1]
    for s in sites[sites["score"] <= thr]:
        p0 = int(s["startPos"]) - firstCol + 1
        for j in range(1, 31):
            p = p0 + j - 1
            if 1 <= p <= L:
                b = int(S[int(s["sequence"]) - 1, p - 1])
                if b:
                    cnt[j, b] += 1
                    cnt[j, 0] += 1
    return cnt.reshape(-1)


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xxmotif_b200 import hostprep, sharded
    nseq, length = 600, 200
    n, codes = hostprep.synthetic_set(nseq, length, seed=5, frac=0.7)
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    pwm, cols = hostprep.profile_pwm()
    half = nseq // world
    lo, hi = rank * half, (rank + 1) * half
    sc, cn, sites, oset = _oracle_shard(n, codes, lo, hi, length, pwm, cols, bg, bgl, rp, rc)
    ex = sharded.Exchange(dist, torch.device("cpu"))
    scs, cns, tot = ex.gather_runs(sc, cn, [len(sites), hi - lo, (hi - lo) * length])
    gsc, gcn = sharded.merge_runs(scs, cns)
    nb, nent, total = tot
    maxSites = total - nent * 7
    K, bp, thr = sharded.choose_subset(gsc, gcn, nb, nent * 1000, maxSites)
    counts = ex.sum_counts(_counts(sites, oset, thr, int(cols[0])))
    new = sharded.build_pwm(counts, 0.1 * K, bg)
    if rank == 0:
        # single-process truth from the oracle on the whole set
        import oracle_lib as ol
        wset = ol.SeqSet(n, codes)
        wnm = ol.NullModel(order=2, rprobs=rp, rconds=rc)
        wctx = ol.Ctx(wset, False, True, False, False, bg, bgl)
        tab = ol.Tables(pwm, cols, nseq, nseq * length, bgl, wnm)
        wctx.scan_full(tab, cols, nseq, (1 - int(cols[0]) + 1, length - int(cols[0]) + 1), True)
        ok, wnb, wbp, wpwm = wctx.update_pwm_with_best_subset(0, pwm, cols, nseq, 0, False, 0.1)
        q.put((nb == wctx.nbSites, K == wnb, bp == wbp, bool((new[5:] == wpwm[5:]).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process():
    import __graft_entry__ as g
    g.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    assert res == (True, True, True, True), res
