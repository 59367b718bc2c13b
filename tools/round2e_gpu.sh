# Final GPU pass of round 2: every GPU test and the full bench line on the code as committed.
R=r02e
mkdir -p gpurun_out
set -x
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "rc $?" >> gpurun_out/${R}_pytest_gpu.log
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 150 python bench.py > gpurun_out/${R}_bench_ours_1gpu.json 2> gpurun_out/${R}_bench_ours_1gpu.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/${R}_bench_ours_1gpu.json"))
    r = d["roofline"]
    print("ms/step %.4f value %.3e scan %.4f frac %.3f e2e %.2f ms refine %.4f s" % (d["ms_per_step"], d["value"], r["kernel_ms"], r["frac"], d["e2e"]["ms_per_step"], d.get("refine_run", {}).get("gpu_seconds", -1)), d["host_phases_us"])
except Exception as e:
    print("no bench line", e)
P
