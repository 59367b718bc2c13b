#!/usr/bin/env python
"""bench.py -- window-positions scored per second of the PWM-refinement hot path on B200 (BASELINE.json metric).

A "step" is one optimizeMotif iteration of the reference (Iterate_Motif.cpp:26-32) over the whole resident sequence
set: reinitialize (p-value tables of the current PWM) -> full scan of every window -> overlap removal -> optimal
subset -> PWM recount; the new PWM feeds the next step.  Workload at N=1: BASELINE.json configs[3] ("-p TATA
profileFile jump-start on 1M x 1 kb sequences"), the configuration the metric is quoted on; it fits one GPU (375 MB
packed).  With --gpus N every rank owns its own 1M x 1 kb shard (weak scaling) and only the score histogram and the
31x5 PWM counts cross NVLink (NCCL all-reduce); the line also carries a strong-scaling leg (1M sequences over all ranks)
and, before anything is timed, a parity check of the N-rank step against the same step on one GPU.

At N=1 the line also carries `configs`: the same step on the other BASELINE configurations and on the shapes the fast path
of round 1 skipped (--revcomp --mops, order-8 negative-set background, --localization with a gapped motif, 12 and 17 match
columns, ragged lengths), each at its BASELINE size and scaled up to fill the GPU, next to a sample run of the reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nseq 1000000] [--seqlen 1000] [--no-matrix]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "window_positions_scored_per_s"
UNIT = "windows/s"
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
PROFILE_TXT = """0.05 0.90 0.05 0.90 0.45 0.90 0.50 0.45
0.05 0.03 0.05 0.03 0.05 0.03 0.05 0.05
0.05 0.03 0.05 0.03 0.05 0.03 0.05 0.45
0.85 0.04 0.85 0.04 0.45 0.04 0.40 0.05
"""


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nseq", type=int, default=1000000)
    ap.add_argument("--seqlen", type=int, default=1000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-seqs", type=int, default=60000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-refine-run", action="store_true")
    ap.add_argument("--no-matrix", action="store_true", help="skip the `configs` array (the other workloads)")
    ap.add_argument("--matrix-only", default="", help="comma-separated workload names for the `configs` array")
    ap.add_argument("--matrix-steps", type=int, default=10)
    ap.add_argument("--refine-sample-seqs", type=int, default=30000)
    ap.add_argument("--no-motif-batch", action="store_true", help="skip the `motif_batch` object (several motifs at once on the small sets)")
    return ap.parse_args()


def workload_name(a):
    return "cfg4: -p TATA profile jump-start refinement step, %d x %d bp per GPU, ZOOPS, bg order 2" % (a.nseq, a.seqlen)


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference (oracle/_ref/ref_harness bench) on host cores
# ---------------------------------------------------------------------------------------------------------------------
def write_fasta(path, n, codes):
    from xxmotif_b200 import workloads
    workloads.write_fasta(path, n, codes)


def run_reference_sample(nseq, seqlen, steps, warmup, nproc):
    """`nproc` independent copies of the (single-threaded) reference on disjoint shards; returns aggregate windows/s."""
    from xxmotif_b200 import hostprep
    tmp = tempfile.mkdtemp(prefix="xxm_bench_ref_")
    prof = os.path.join(tmp, "tata.prof")
    with open(prof, "w") as fh:
        fh.write(PROFILE_TXT)
    procs = []
    for p in range(nproc):
        n, codes = hostprep.synthetic_set(nseq, seqlen, seed=1000 + p)
        fa = os.path.join(tmp, "in%d.fa" % p)
        write_fasta(fa, n, codes)
        cmd = [HARNESS, "bench", str(steps), str(warmup), os.path.join(tmp, "out%d" % p), fa, "--batch", "-p", prof]
        procs.append(cmd)
    t0 = time.time()
    running = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL) for c in procs]
    outs = [r.communicate()[0].decode() for r in running]
    wall = time.time() - t0
    tot = 0.0
    ms = []
    for o in outs:
        line = [x for x in o.splitlines() if x.startswith("{")]
        if not line:
            raise RuntimeError("reference harness produced no result")
        j = json.loads(line[-1])
        tot += j["windows_per_s"]
        ms.append(1000.0 * j["seconds"] / max(1, j["steps"]))
    subprocess.call(["rm", "-rf", tmp])
    return tot, sum(ms) / len(ms), wall


def run_reference_workload(w, steps, warmup, n=None, codes=None, neg=None):
    """the reference's optimizeMotif loop body on one workload of xxmotif_b200.workloads (one process, one core): its JSON line"""
    from xxmotif_b200 import workloads
    tmp = tempfile.mkdtemp(prefix="xxm_bench_wl_")
    try:
        if n is None:
            n, codes = w.positive_set()
            neg = w.negative_set(n, codes)
        fa = os.path.join(tmp, "in.fa")
        workloads.write_fasta(fa, n, codes)
        negfa = None
        if neg is not None:
            negfa = os.path.join(tmp, "neg.fa")
            workloads.write_fasta(negfa, neg[0], neg[1])
        cmd = [HARNESS, "bench", str(steps), str(warmup), os.path.join(tmp, "out"), fa] + w.reference_flags(tmp, negfa)
        o = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        line = [x for x in o.splitlines() if x.startswith("{")]
        if not line:
            raise RuntimeError("reference harness produced no result for %s" % w.name)
        return json.loads(line[-1])
    finally:
        subprocess.call(["rm", "-rf", tmp])


def run_reference_refine(nseq, seqlen):
    """The unmodified reference's MotifContainer::iterate of the jump-start motif (ref_harness refine) on an nseq sample."""
    from xxmotif_b200 import hostprep
    tmp = tempfile.mkdtemp(prefix="xxm_bench_refine_")
    try:
        prof = os.path.join(tmp, "tata.prof")
        with open(prof, "w") as fh:
            fh.write(PROFILE_TXT)
        n, codes = hostprep.synthetic_set(nseq, seqlen, seed=4)
        fa = os.path.join(tmp, "in.fa")
        write_fasta(fa, n, codes)
        o = subprocess.run([HARNESS, "refine", os.path.join(tmp, "out"), fa, "--batch", "-p", prof], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        line = [x for x in o.splitlines() if x.startswith("{")]
        if not line:
            raise RuntimeError("reference harness produced no result")
        return json.loads(line[-1])
    finally:
        subprocess.call(["rm", "-rf", tmp])


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(HARNESS):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness not built (make -C oracle ref)"}))
        return 0
    ncores = os.cpu_count() or 1
    nproc = max(1, min(ncores, 32))
    sample = min(a.nseq, 10000)
    steps = max(1, min(a.steps, 5))
    warmup = max(1, min(a.warmup, 2))
    v, ms, wall = run_reference_sample(sample, a.seqlen, steps, warmup, nproc)
    out = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(a), "inputs": "reference CPU path, no GPU"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nproc, "kind": "reference",
                         "sample": "%d independent single-threaded reference processes x (%d x %d bp shard), %d steps each; host has %d cores"
                                   % (nproc, sample, a.seqlen, steps, ncores)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region: NVML polled from a thread every ~2 ms (nvidia-smi -lms needs tens of
# milliseconds per sample and missed a 130 ms timed region almost entirely)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.rows = []
        self.stop_flag = False
        self.th = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.h, self.nv = h, pynvml
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception as ex:
            self.err = str(ex)

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception:
                pass
            time.sleep(0.002)

    def clear(self):
        self.rows = []

    def stop(self):
        if not self.th:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: %s" % self.err], "samples": 0}
        self.stop_flag = True
        self.th.join(timeout=2)
        rows = list(self.rows)
        sm = sorted(r[0] for r in rows)
        reasons = set()
        for _, bits in rows:
            for b, nm in self.REASONS.items():
                if bits & b:
                    reasons.add(nm)
        return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.mx), "reasons": sorted(reasons), "samples": len(sm),
                "how": "NVML polled every ~2 ms during the timed region"}


def windows_of(n, span, revcomp):
    """window positions the reference's scan loop visits (StartPosUpdater-inl.h:90-96): stored length - span + 1 per sequence"""
    import numpy as np
    stored = 2 * n.astype(np.int64) + 1 if revcomp else n.astype(np.int64)
    return int((stored - span + 1).clip(min=0).sum()), int(stored.sum())


# ---------------------------------------------------------------------------------------------------------------------
# the `configs` array: the same step on the other workloads
# ---------------------------------------------------------------------------------------------------------------------
def time_workload(ctx, w, n, codes, neg, ref, steps, peak):
    """load workload `w` into ctx and time `steps` optimizeMotif steps from the reference's start motif"""
    import numpy as np
    from xxmotif_b200 import hostprep, localization
    ctx.set_options(mops=w.mops, oops=False, localization=w.localization, bg=[0, .25, .25, .25, .25], bgLog=[0] + [float(np.log(.25))] * 4)
    t0 = time.perf_counter()
    if neg is not None:  # the background comes from the negative set: count it on the device, then load the positive set
        ctx.load_seqset(neg[0], neg[1], revcomp=w.revcomp)
        bg, bgl = hostprep.background_device(ctx)
        rp, rc = hostprep.null_model_device(ctx, order=w.order)
        ctx.load_seqset(n, codes, revcomp=w.revcomp)
    else:
        ctx.load_seqset(n, codes, revcomp=w.revcomp)
        bg, bgl = hostprep.background_device(ctx)
        rp, rc = hostprep.null_model_device(ctx, order=w.order)
    ctx.set_options(mops=w.mops, oops=False, localization=w.localization, bg=bg, bgLog=bgl)
    ctx.set_background(w.order, rp, rc)
    setup_s = time.perf_counter() - t0
    total, _, mxl, _ = ctx.seqset_info()
    nent = len(n)
    cols = np.asarray(ref["cols"], dtype=np.int32)
    c0, span = int(cols[0]), int(cols[-1] - cols[0] + 1)
    pwm0 = np.asarray(ref["pwm0"], dtype=np.float64)
    raw_max = int(n.max())
    enr0 = (1 - c0, 1 - c0, raw_max - c0 + 1)  # Global::startRegion/endRegion defaults over the sequences as read (Motif.cpp:791-793)
    windows, stored = windows_of(n, span, w.revcomp)

    def run(k, record):
        pwm, enr = pwm0.copy(), enr0
        scan_ms, hits, info = [], [], None
        for _ in range(k):
            pwm, info = ctx.optimize_step(pwm, cols, nent, int(total[nent]), (enr[1], enr[2]), 0.1)
            if w.localization:
                enr = localization.update_enriched_region(ctx, c0, span, enr)
            if record:
                st = ctx.stats()
                scan_ms.append(st["last_scan_ms"])
                hits.append(st["last_scan_hits"])
        return info, scan_ms, hits

    run(3, False)
    ctx.step_phases(reset=True)
    if os.environ.get("XXM_TRACE"):  # per-call wall times of the timed steps only (the dump clears what the warm-up collected)
        sys.stderr.write("---- warm-up calls\n")
        sys.stderr.flush()
        ctx.L.xxm_trace_dump()
        sys.stderr.write("---- timed steps\n")
        sys.stderr.flush()
    import torch
    torch.cuda.synchronize()
    l0 = ctx.stats()["launches"]
    t0 = time.perf_counter()
    info, scan_ms, hits = run(steps, True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1000.0 / steps
    launches = ctx.stats()["launches"] - l0
    phases, _ = ctx.step_phases(reset=True)
    sm = sum(scan_ms) / len(scan_ms)
    hb = sum(hits) / len(hits)
    alg = 0.375 * stored + 12.0 * hb
    return {"windows_per_step": windows, "stored_bases": stored, "ms_per_step": ms, "value": windows / (ms / 1000.0), "unit": UNIT,
            "match_columns": int(len(cols)), "hits_per_step": int(hb), "sites": int(info["nbSites"]), "bestMotifNb": int(info["bestMotifNb"]),
            "scan": {"kernel_ms": sm, "achieved_GBs": alg / (sm / 1000.0) / 1e9 if sm > 0 else None, "frac": (alg / (sm / 1000.0) / 1e9 / peak) if sm > 0 else None,
                     "windows_per_s": windows / (sm / 1000.0) if sm > 0 else None},
            "host_phases_us": {k: round(v, 1) for k, v in phases.items()}, "gpu_launches_per_step": launches / steps, "setup_s": round(setup_s, 2)}


def matrix_lines(ctx, a, peak, main_entry):
    import numpy as np
    from xxmotif_b200 import workloads as W
    only = set(x for x in a.matrix_only.split(",") if x)
    out = []
    for w, scale in W.MATRIX:
        if only and w.name not in only:
            continue
        e = {"name": w.name, "what": w.what}
        try:
            if w.name == "cfg4":
                e.update({"at_scale": dict(main_entry, workload=workload_name(a))})
                out.append(e)
                continue
            # the reference on a sample of the shape (one core): start motif + its windows/s
            sample_n = 1500 if w.name == "cols17" else 5000
            ws = w.scaled(sample_n, neg=20000 if w.neg else None)
            t0 = time.perf_counter()
            ref = run_reference_workload(ws, 2, 1)
            e["reference_sample"] = {"windows_per_s": ref["windows_per_s"], "cores": 1, "sequences": sample_n, "steps": 2,
                                     "seconds": round(time.perf_counter() - t0, 1), "workload": ws.describe()}
            for label, nseq in (("baseline_size", w.nseq), ("at_scale", scale)):
                if label == "at_scale" and scale == w.nseq:
                    e[label] = "same as baseline_size"
                    continue
                wl = w.scaled(nseq)
                n, codes = wl.positive_set()
                neg = wl.negative_set(n, codes)
                r = time_workload(ctx, wl, n, codes, neg, ref, a.matrix_steps, peak)
                r["workload"] = wl.describe()
                r["vs_reference_1core"] = r["value"] / ref["windows_per_s"]
                e[label] = r
                del n, codes, neg
        except Exception as ex:  # keep the line
            e["failed"] = "%s: %s" % (type(ex).__name__, ex)
        out.append(e)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# N > 1: the sharded step must return what one GPU returns on the union of the shards
# ---------------------------------------------------------------------------------------------------------------------
def parity_nranks(ctx, dist, rank, world, local):
    """N x 5000 x 500 bp: three steps over NCCL on N ranks vs the same three steps on rank 0 alone over the union"""
    import numpy as np
    import torch
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded
    per, L = 5000, 500
    n_all, codes_all = hostprep.synthetic_set(per * world, L, seed=77)
    bg, bgl = hostprep.background(n_all, codes_all)
    rp, rc = hostprep.null_model(n_all, codes_all, order=2)
    pwm0, cols = hostprep.profile_pwm()
    enr = (1 - int(cols[0]) + 1, L - int(cols[0]) + 1)
    lo, hi = rank * per, (rank + 1) * per
    ctx.set_options(mops=False, oops=False, localization=False, bg=bg, bgLog=bgl)
    ctx.set_background(2, rp, rc)
    ctx.load_seqset(n_all[lo:hi], codes_all[lo * L:hi * L], revcomp=False)
    job = sharded.ShardedRefinement(ctx, dist, nent_local=per, total_local=per * L, pseudo=0.1, bg=bg, comm_ready=True)
    pwm = pwm0.copy()
    infos = []
    for _ in range(3):
        pwm, info = job.optimize_step(pwm, cols, enr)
        infos.append((int(info["nbSites_global"]), int(info["bestMotifNb"]), float(info["bestPval"])))
    res = {"ok": True, "ranks": world, "sequences": per * world, "steps": 3}
    flag = torch.zeros(1, device="cuda", dtype=torch.int32)
    if rank == 0:
        solo = xb.Context(local)
        try:
            solo.set_options(mops=False, oops=False, localization=False, bg=bg, bgLog=bgl)
            solo.set_background(2, rp, rc)
            solo.load_seqset(n_all, codes_all, revcomp=False)
            p1 = pwm0.copy()
            ref = []
            for _ in range(3):
                p1, i1 = solo.optimize_step(p1, cols, per * world, per * world * L, enr, 0.1)
                ref.append((int(i1["nbSites"]), int(i1["bestMotifNb"]), float(i1["bestPval"])))
        finally:
            solo.close()
        same = ref == infos and bool((p1 == pwm).all())
        res.update({"ok": bool(same), "nbSites_bestMotifNb_bestPval": infos[-1], "single_gpu": ref[-1], "pwm_identical": bool((p1 == pwm).all())})
        flag[0] = 0 if same else 1
    dist.broadcast(flag, src=0)
    res["ok"] = bool(int(flag.item()) == 0)
    return res


def tool_json(argv, timeout):
    """Run one of tools/*.py in a process of its own and return the last JSON object it printed, or {"failed": why}: a side
    measurement that crashes, hangs or prints nothing must not cost the bench line its numbers."""
    try:
        r = subprocess.run([sys.executable] + list(argv), capture_output=True, text=True, timeout=timeout)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if not lines:
            return {"failed": "rc %d: %s" % (r.returncode, (r.stderr or "").strip()[-300:])}
        res = json.loads(lines[-1])
        if r.returncode:
            res["rc"] = r.returncode
        return res
    except Exception as ex:
        return {"failed": "%s: %s" % (type(ex).__name__, ex)}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def ours(a):
    # rank 0 prints ONE JSON line on stdout.  NCCL (NCCL_DEBUG=INFO/VERSION) and anything else that writes to fd 1 goes to
    # stderr instead: fd 1 is pointed at fd 2 for the whole run and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out, rank = _ours(a)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    if rank == 0:
        os.write(1, (json.dumps(out) + "\n").encode())
    return 0 if out.get("ok", True) else 3


def _ours(a):
    import numpy as np
    import torch
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    # packing threads of xxm_load_seqset (the e2e leg): the ranks of one box share its cores
    # (a rank with 12 or more cores to itself leaves two of them to its own thread and the driver's: 14 packers on the 16-core box
    # load 1 GB in 14.0-14.5 ms, 12 in 14.6-14.8 ms, 16 in 15.1 ms -- profiles/r02b_load_paths_hybrid_threads.txt, r02_load_paths.json)
    per_rank = (os.cpu_count() or 1) // max(1, world)
    pack_threads = max(1, min(14, per_rank - 2)) if per_rank >= 12 else max(1, (3 * per_rank) // 4)
    os.environ.setdefault("XXM_PACK_THREADS", str(pack_threads))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    ctx = xb.Context(local)
    parity = None
    if dist:
        # one NCCL communicator for the library (its own, on its own stream), created once
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.from_numpy(xb.capi.comm_unique_id()))
        dist.broadcast(idt, src=0)
        ctx.comm_init(rank, world, idt.cpu().numpy())
        parity = parity_nranks(ctx, dist, rank, world, local)
        if not parity["ok"]:
            ctx.close()
            dist.destroy_process_group()
            return {"metric": METRIC, "ok": False, "parity_nranks": parity, "error": "the N-rank step differs from the single-GPU step"}, rank

    # ---- synthetic shard of this rank, background, NullModel (inputs of the path)
    n, codes = hostprep.synthetic_set(a.nseq, a.seqlen, seed=4 + rank)
    pinned = torch.empty(codes.size, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = codes
    codes = pinned.numpy()
    bg, bgl = hostprep.background(n, codes)
    rp, rc = hostprep.null_model(n, codes, order=2)
    if dist:  # one background model for the whole set (rank 0's shard stands in for it): every rank must build the same tables
        for arr in (bg, bgl, rp, rc):
            t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
            dist.broadcast(t, src=0)
            arr[...] = t.cpu().numpy()
    ctx.set_options(mops=False, oops=False, localization=False, bg=bg, bgLog=bgl)
    ctx.set_background(2, rp, rc)
    ctx.load_seqset(n, codes, revcomp=False)
    total, avg, mxl, mnl = ctx.seqset_info()
    pwm0, cols = hostprep.profile_pwm()
    nent = len(n)
    span = int(cols[-1] - cols[0] + 1)
    windows_local, stored_local = windows_of(n, span, False)
    enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)  # Global::startRegion/endRegion defaults, Motif.cpp:791-793
    job = sharded.ShardedRefinement(ctx, dist, nent_local=nent, total_local=int(total[nent]), pseudo=0.1, bg=bg, comm_ready=True)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def step(pwm):
        return job.optimize_step(pwm, cols, enr)

    def timed_steps(k, pwm):
        """k steps bracketed by barrier + synchronize; returns (ms max over ranks, device ms, wall ms, per-step scan records, pwm, info)"""
        rec = []
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        info = None
        for _ in range(k):
            pwm, info = step(pwm)
            st = ctx.stats()
            rec.append((st["last_scan_ms"], st["last_find_ms"], st["last_scan_hits"]))
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1000.0
        if dist:
            dist.barrier()
        dev = e0.elapsed_time(e1)
        ms = max(dev, wall) if abs(dev - wall) > 0.2 * wall else dev
        if dist:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, dev, wall, rec, pwm, info

    pwm = pwm0.copy()
    for _ in range(max(a.warmup, 3)):
        pwm, info = step(pwm)
    uuid = None
    try:
        uuid = torch.cuda.get_device_properties(local).uuid
    except Exception:
        pass
    sampler = ClockSampler(local, uuid)
    sampler.start()
    for _ in range(3):  # keep the GPU busy while the sampler starts
        pwm, info = step(pwm)
    ctx.step_phases(reset=True)
    sampler.clear()
    launches0 = ctx.stats()["launches"]
    # ---- timed region: K steps, barrier + synchronize on both sides, CUDA events on the library's stream
    ms_total, dev_ms, wall_ms, rec, pwm, info = timed_steps(a.steps, pwm)
    clocks = sampler.stop()
    launches = ctx.stats()["launches"] - launches0
    phases, _ = ctx.step_phases(reset=True)
    ms_per_step = ms_total / a.steps
    value = windows_local * world * a.steps / (ms_total / 1000.0)

    # ---- e2e: nothing resident.  Every step uploads the raw sequence set from pinned host memory through the C ABI,
    # packs it, runs the step and reads the new PWM back.
    e2e_steps = max(1, a.e2e_steps)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p2 = pwm0.copy()
    for _ in range(e2e_steps):
        ctx.load_seqset(n, codes, revcomp=False)
        p2, _ = step(p2)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    if dist:
        tt = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = windows_local * world * e2e_steps / (e2e_ms / 1000.0)
    packed_words = int(((n.astype(np.int64) + 32 + 31) // 32).sum()) + 2
    host_pack = int(os.environ["XXM_PACK_THREADS"]) >= 6   # xxm_load_seqset packs on the host cores when it has that many, else on the device
    raw_chunks, all_chunks = ctx.last_load_info()
    if all_chunks:  # 2-5 host threads per GPU: both at once
        fr = raw_chunks / all_chunks
        h2d = int(fr * codes.size + (1 - fr) * 12 * packed_words + 4 * nent + 64 * 1024)
        d2h = int(job.last_d2h_bytes + fr * 4 * nent)
        how = ("%s host threads pack 8 MB chunks from the back of the set while raw chunks from the front cross the bus and are packed on the device: "
               "%d of %d chunks went raw in the last step" % (os.environ["XXM_PACK_THREADS"], raw_chunks, all_chunks))
    elif host_pack:
        h2d = int(12 * packed_words + 4 * nent + 64 * 1024)   # packed stream (8 B codes + 4 B N mask per 32 bases) + lengths + the per-model tables of the step
        d2h = int(job.last_d2h_bytes)                          # step results
        how = "%s host threads pack 4 MB chunks into the 2-bit layout, each chunk copied as soon as it is packed" % os.environ["XXM_PACK_THREADS"]
    else:
        h2d = int(codes.size + 4 * nent + 64 * 1024)           # raw codes + lengths + tables
        d2h = int(job.last_d2h_bytes + 4 * nent)               # step results + per-sequence residue counts
        how = "one host core per GPU: raw bytes copied in chunks, statistics and packing on the device behind them"

    # ---- roofline of the dominant kernel (full scan): algorithmic bytes / measured kernel time.  Bytes per SURVEY 8(d):
    # 0.375 B per base of the sequences (sum of n_i -- the padding between sequences is not counted) + 12 B per hit
    scan_ms = [r[0] for r in rec]
    find_ms = [r[1] for r in rec]
    hits = [r[2] for r in rec]
    avg_ms = sum(scan_ms) / len(scan_ms)
    alg_bytes = 0.375 * stored_local + 12.0 * (sum(hits) / len(hits))
    achieved = alg_bytes / (avg_ms / 1000.0) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "scan_traffic_bytes.json")
    if os.path.exists(tpath) and a.nseq == 1000000 and a.seqlen == 1000:  # the capture belongs to the default workload
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "full scan = k_scan5 (find, dead windows cleared, per-chunk counts) + k_tile_scan + %s (ordered site records)" % ("k_expand6" if os.environ.get("XXM_SCAN_GEN") == "6" else "k_expand7"), "kernel_ms": avg_ms,
                "find_ms": sum(find_ms) / len(find_ms), "expand_ms": avg_ms - sum(find_ms) / len(find_ms), "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes": "0.375 B per base (sum of sequence lengths) + 12 B per hit = %.1f MB per scan" % (alg_bytes / 1e6), "step_share": avg_ms / ms_per_step,
                # the same figure against the nominal 8 TB/s, and in the reference's own representation (1 B per base)
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "achieved_at_1B_per_base": (1.0 * stored_local + 12.0 * (sum(hits) / len(hits))) / (avg_ms / 1000.0) / 1e9}

    main_entry = {"windows_per_step": windows_local, "stored_bases": stored_local, "ms_per_step": ms_per_step, "value": value / world, "unit": UNIT, "match_columns": int(len(cols)),
                  "hits_per_step": int(info["nbBestSites"]), "sites": int(info["nbSites"]), "bestMotifNb": int(info["bestMotifNb"]),
                  "scan": {"kernel_ms": avg_ms, "achieved_GBs": achieved, "frac": achieved / peak, "windows_per_s": windows_local / (avg_ms / 1000.0)},
                  "host_phases_us": {k: round(v, 1) for k, v in phases.items()}, "gpu_launches_per_step": launches / a.steps}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "windows_per_step_per_gpu": windows_local, "hits_per_step_per_gpu": int(info["nbBestSites"]),
                   "sites_after_overlap_removal": int(info["nbSites"]), "bestMotifNb": int(info["bestMotifNb"]),
                   "l2": "inputs larger than L2 (packed stream %.0f MB per GPU)" % (0.375 * nent * (a.seqlen + 32 + 24) / 1e6),
                   "timing": "CUDA events on the library stream, max over ranks; device %.3f ms vs wall %.3f ms" % (dev_ms, wall_ms)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "what": "every step: xxm_load_seqset from the reference's byte-per-base host arrays (%s) + xxm_optimize_step (PWM in and out through host memory)" % how, "host_bytes_read_per_step": int(codes.size) if host_pack else 0},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "host_phases_us": {k: round(v, 1) for k, v in phases.items()},
    }
    if parity is not None:
        out["parity_nranks"] = parity
    # ---- strong scaling: the 1M-sequence set of the metric spread over all ranks (a.nseq / N sequences each)
    if dist:
        try:
            ns = max(1, a.nseq // world)
            ctx.load_seqset(n[:ns], codes[:ns * a.seqlen], revcomp=False)
            tot_s, _, _, _ = ctx.seqset_info()
            job_s = sharded.ShardedRefinement(ctx, dist, nent_local=ns, total_local=int(tot_s[ns]), pseudo=0.1, bg=bg, comm_ready=True)
            job_main, job = job, job_s
            p3 = pwm0.copy()
            for _ in range(5):
                p3, _ = step(p3)
            ctx.step_phases(reset=True)
            k = max(20, a.steps // 2)
            ms_s, _, _, rec_s, p3, info_s = timed_steps(k, p3)
            ph_s, _ = ctx.step_phases(reset=True)
            w_s, _ = windows_of(n[:ns], span, False)
            out["strong_1M"] = {"sequences_per_gpu": ns, "ms_per_step": ms_s / k, "value": w_s * world * k / (ms_s / 1000.0), "unit": UNIT, "steps": k,
                                "scan_ms": sum(r[0] for r in rec_s) / len(rec_s), "host_phases_us": {kk: round(v, 1) for kk, v in ph_s.items()}}
            job = job_main
        except Exception as ex:
            out["strong_1M"] = {"failed": str(ex)}
    # ---- "end-to-end s" of the metric: the whole refinement of the jump-start motif (MotifContainer::iterate's per-motif loop:
    # optimizeMotif + maximizeMotifLength until the p-value stops improving) on the resident set, one xxm_iterate_motif call
    if world == 1 and not a.no_refine_run:
        try:
            from xxmotif_b200 import refine
            _, _, mxl, _ = ctx.seqset_info()
            c0 = int(cols[0])
            mot = refine.Motif(pwm0, cols, nent, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=max(1 << 20, 2 * nent))
            R = refine.Refiner(ctx, totalPositions=int(total[nent]), revcomp=False, localization=False, maximizeMotifLength=True, minMatchPositions=5,
                               maxMatchPositions=17, pseudo=0.1, neff_pwm=10.0)
            secs = []
            for _ in range(2):  # the first run on a context also pays for buffers it allocates (a real run refines hundreds of motifs on one context)
                mot = refine.Motif(pwm0, cols, nent, enr=(1 - c0, 1 - c0, int(mxl) - c0 + 1), siteCap=max(1 << 20, 2 * nent))
                torch.cuda.synchronize()
                l0 = ctx.stats()["launches"]
                t0 = time.perf_counter()
                rounds = R.iterate_motif(mot)
                torch.cuda.synchronize()
                secs.append(time.perf_counter() - t0)
            gpu_s = secs[-1]
            rr = {"gpu_seconds": gpu_s, "gpu_seconds_first_run_on_context": secs[0], "sequences": nent, "rounds": int(rounds),
                  "match_columns": [int(mot.cols[0]), int(mot.cols[-1])],
                  "sites": int(mot.m.nsites), "log_pval": float(mot.pval), "gpu_launches": int(ctx.stats()["launches"] - l0),
                  "what": "xxm_iterate_motif of the TATA profile on the resident set (set load not included)"}
            if not a.no_cpu_baseline and os.path.exists(HARNESS):
                ns = min(a.refine_sample_seqs, a.nseq)
                rj = run_reference_refine(ns, a.seqlen)
                rr.update({"reference_seconds_sample": rj["seconds"], "reference_sample_sequences": ns,
                           "reference_seconds_scaled": rj["seconds"] * nent / ns, "reference_cores": 1,
                           "reference_sample_result": {"match_columns": [rj["firstCol"], rj["lastCol"]], "sites": rj["nsites"], "log_pval": rj["pval"]}})
            out["refine_run"] = rr
        except Exception as ex:
            out["refine_run"] = {"failed": str(ex)}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            if os.path.exists(HARNESS):
                v, ms, wall = run_reference_sample(min(a.cpu_sample_seqs, a.nseq), a.seqlen, 3, 1, 1)
                out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                                       "sample": "unmodified reference (oracle/_ref/ref_harness), %d x %d bp of the same workload, 3 steps, %.1f s wall"
                                                 % (min(a.cpu_sample_seqs, a.nseq), a.seqlen, wall)}
            else:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref/ref_harness missing"}
        except Exception as ex:  # keep the GPU line even if the host sample fails
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "failed: %s" % ex}
    if world == 1 and not a.no_matrix and os.path.exists(HARNESS):
        del codes, pinned
        out["configs"] = matrix_lines(ctx, a, peak, main_entry)
    ctx.close()
    # ---- SURVEY 8(f) row 2, "batch motifs": 16 start models refined one after the other (the reference's loop) and four at a time
    # on views of one resident set, on the small sets of configs 1 and 2.  A separate process with its own limit (tools/motif_batch.py):
    # whatever happens there cannot cost the line above its numbers.
    if world == 1 and not a.no_motif_batch:
        out["motif_batch"] = tool_json([os.path.join(ROOT, "tools", "motif_batch.py"), "--shapes", "cfg1,cfg2", "--contexts", "4", "--motifs", "16"], 120)
    # ---- the other half of that row: the column variants of maximizeMotifLength of the ONE jump-start motif spread over the
    # context and its views (xxm_iterate_motif_on), on the bench's own set size -- again in a process of its own
    if world == 1 and not a.no_motif_batch and not a.no_refine_run:
        vb = tool_json([os.path.join(ROOT, "tools", "variant_batch.py"), "--sizes", str(a.nseq), "--seqlen", str(a.seqlen), "--contexts", "1,2,4"], 180)
        if isinstance(out.get("refine_run"), dict):
            out["refine_run"]["variants_on_views"] = vb
        else:
            out["variants_on_views"] = vb
    if dist:
        dist.destroy_process_group()
    return out, rank


if __name__ == "__main__":
    args = parse()
    sys.exit(reference_arm(args) if args.impl == "reference" else ours(args))
