# Third GPU pass of round 2: the fused re-score kernel (k_eval_select_zoops) and the ragged origins of k_expand7.
R=r02c
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "rc $?" >> gpurun_out/${R}_pytest_gpu.log
tail -3 gpurun_out/${R}_pytest_gpu.log
for f in 1 0; do
  if [ "$f" = "1" ]; then export XXM_NO_FUSED_RESCORE=1; else unset XXM_NO_FUSED_RESCORE; fi
  timeout 200 python bench.py --steps 50 --no-cpu-baseline --matrix-only ragged --matrix-steps 30 --e2e-steps 1 > gpurun_out/${R}_bench_nofused$f.json 2> gpurun_out/${R}_bench_nofused$f.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${R}_bench_nofused$f.json"))
    rr = d.get("refine_run", {})
    print("nofused=$f ms_per_step", d["ms_per_step"], "refine_run", rr.get("gpu_seconds"), rr.get("gpu_seconds_first_run_on_context"), rr.get("sites"), rr.get("log_pval"))
    for c in d.get("configs", []):
        for k in ("at_scale", "baseline_size"):
            if k in c: print(" ", c["name"], k, c[k].get("ms_per_step"), c[k].get("scan"))
except Exception as e:
    print("no bench line", e)
P
done
XXM_TRACE=1 timeout 120 python tools/refine_trace.py > gpurun_out/${R}_refine_trace.txt 2>&1; tail -25 gpurun_out/${R}_refine_trace.txt
