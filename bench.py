#!/usr/bin/env python
"""bench.py -- window-positions scored per second of the PWM-refinement hot path on B200 (BASELINE.json metric).

A "step" is one optimizeMotif iteration of the reference (Iterate_Motif.cpp:26-32) over the whole resident sequence
set: reinitialize (p-value tables of the current PWM) -> full scan of every window -> overlap removal -> optimal
subset -> PWM recount; the new PWM feeds the next step.  Workload at N=1: BASELINE.json configs[3] ("-p TATA
profileFile jump-start on 1M x 1 kb sequences"), the configuration the metric is quoted on; it fits one GPU (375 MB
packed).  With --gpus N every rank owns its own 1M x 1 kb shard (weak scaling) and only the score histogram and the
31x5 PWM counts cross NVLink (NCCL all-gather / all-reduce).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nseq 1000000] [--seqlen 1000]
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
    ap.add_argument("--refine-sample-seqs", type=int, default=30000)
    return ap.parse_args()


def workload_name(a):
    return "cfg4: -p TATA profile jump-start refinement step, %d x %d bp per GPU, ZOOPS, bg order 2" % (a.nseq, a.seqlen)


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference (oracle/_ref/ref_harness bench) on host cores
# ---------------------------------------------------------------------------------------------------------------------
def write_fasta(path, n, codes):
    import numpy as np
    lut = np.frombuffer(b"NACGT", dtype=np.uint8)
    L = int(n[0])
    rows = lut[codes.reshape(len(n), L)]
    with open(path, "wb") as fh:
        for i in range(len(n)):
            fh.write(b">seq%d\n" % (i + 1))
            fh.write(rows[i].tobytes())
            fh.write(b"\n")


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
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.decode().strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def ours(a):
    import numpy as np
    import torch
    import xxmotif_b200 as xb
    from xxmotif_b200 import hostprep, sharded

    # rank 0 prints ONE JSON line on stdout: NCCL writes its version banner there from level VERSION up (WARN included)
    if os.environ.get("XXM_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["XXM_NCCL_DEBUG"]
    else:
        os.environ.pop("NCCL_DEBUG", None)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # packing threads of xxm_load_seqset (the e2e leg): the ranks of one box share its cores
    pack_threads = max(1, min(12, (3 * (os.cpu_count() or 1)) // (4 * max(1, world))))
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
    ctx = xb.Context(local)
    ctx.set_options(mops=False, oops=False, localization=False, bg=bg, bgLog=bgl)
    ctx.set_background(2, rp, rc)
    ctx.load_seqset(n, codes, revcomp=False)
    total, avg, mxl, mnl = ctx.seqset_info()
    pwm0, cols = hostprep.profile_pwm()
    nent = len(n)
    span = int(cols[-1] - cols[0] + 1)
    windows_local = int((n.astype(np.int64) - span + 1).clip(min=0).sum())
    enr = (1 - int(cols[0]) + 1, mxl - int(cols[0]) + 1)  # Global::startRegion/endRegion defaults, Motif.cpp:791-793
    job = sharded.ShardedRefinement(ctx, dist, nent_local=nent, total_local=int(total[nent]), pseudo=0.1, bg=bg)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def step(pwm):
        return job.optimize_step(pwm, cols, enr)

    pwm = pwm0.copy()
    for _ in range(max(a.warmup, 3)):
        pwm, info = step(pwm)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(3):  # keep the GPU busy while nvidia-smi starts sampling
        pwm, info = step(pwm)
    # ---- timed region: K steps, barrier + synchronize on both sides, CUDA events on the library's stream
    scan_ms, find_ms, scan_bytes, scan_bytes_ref = [], [], [], []
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.rows.clear()
    launches0 = ctx.stats()["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        pwm, info = step(pwm)
        st = ctx.stats()
        scan_ms.append(st["last_scan_ms"])
        find_ms.append(st["last_find_ms"])
        scan_bytes.append(0.375 * st["last_scan_bases"] + 12.0 * st["last_scan_hits"])
        scan_bytes_ref.append(1.0 * st["last_scan_bases"] + 12.0 * st["last_scan_hits"])
    e1.record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = ctx.stats()["launches"] - launches0
    ms_total = max(dev_ms, wall_ms) if abs(dev_ms - wall_ms) > 0.2 * wall_ms else dev_ms
    if dist:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
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
    if host_pack:
        h2d = int(12 * packed_words + 4 * nent + 64 * 1024)   # packed stream (8 B codes + 4 B N mask per 32 bases) + lengths + the per-model tables of the step
        d2h = int(job.last_d2h_bytes)                          # step results
        how = "%s host threads pack 8 MB chunks into the 2-bit layout, each chunk copied as soon as it is packed" % os.environ["XXM_PACK_THREADS"]
    else:
        h2d = int(codes.size + 4 * nent + 64 * 1024)           # raw codes + lengths + tables
        d2h = int(job.last_d2h_bytes + 4 * nent)               # step results + per-sequence residue counts
        how = "fewer than 6 host cores per GPU: raw bytes copied in chunks, statistics and packing on the device behind them"

    # ---- roofline of the dominant kernel (full scan): algorithmic bytes / measured kernel time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_ms = sum(scan_ms) / len(scan_ms)
    achieved = (sum(scan_bytes) / len(scan_bytes)) / (avg_ms / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scan_traffic_bytes.json")
    if os.path.exists(tpath) and a.nseq == 1000000 and a.seqlen == 1000:  # the capture belongs to the default workload
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "full scan = k_scan5 (find, dead windows cleared, per-chunk counts) + k_tile_scan + k_expand6 (ordered site records)", "kernel_ms": avg_ms,
                "find_ms": sum(find_ms) / len(find_ms), "expand_ms": avg_ms - sum(find_ms) / len(find_ms), "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes": "0.375 B per stored base + 12 B per hit", "step_share": avg_ms / ms_per_step,
                # the same figure against the nominal 8 TB/s, and in the reference's own representation (1 B per base)
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "achieved_at_1B_per_base": (sum(scan_bytes_ref) / len(scan_bytes_ref)) / (avg_ms / 1000.0) / 1e9}

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
    }
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
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rounds = R.iterate_motif(mot)
            torch.cuda.synchronize()
            gpu_s = time.perf_counter() - t0
            rr = {"gpu_seconds": gpu_s, "sequences": nent, "rounds": int(rounds), "match_columns": [int(mot.cols[0]), int(mot.cols[-1])],
                  "sites": int(mot.m.nsites), "log_pval": float(mot.pval),
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
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(reference_arm(args) if args.impl == "reference" else ours(args))
