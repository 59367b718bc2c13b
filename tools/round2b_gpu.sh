# Second GPU pass of round 2 (gpurun -- bash tools/round2b_gpu.sh): the generation-7 scan kernels.  Ordered by priority; every
# step has its own timeout; each ncu pass runs after the same command has exited 0 without ncu.
R=r02b
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${R}_gpu.txt 2>&1
# 1. every GPU test on the default kernels (generation 7)
timeout 600 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "rc $?" >> gpurun_out/${R}_pytest_gpu.log
tail -4 gpurun_out/${R}_pytest_gpu.log
# 2. A/B on one box: the bench step with the previous and with the new scan kernels
for g in 6 7; do
  XXM_SCAN_GEN=$g timeout 200 python bench.py --steps 100 --no-cpu-baseline --no-refine-run --no-matrix --e2e-steps 1 > gpurun_out/${R}_bench_gen$g.json 2> gpurun_out/${R}_bench_gen$g.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/${R}_bench_gen$g.json"))
    print("gen $g ms_per_step", d["ms_per_step"], "scan", d["roofline"]["kernel_ms"], "find", d["roofline"]["find_ms"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("gen $g: no bench line", e)
P
done
# 3. launch list and full capture of the step kernels (generation 7)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-matrix --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_scan5|k_tile_scan|k_expand7|k_select_cls_fwd|k_count_pwm3" -s 10 -c 5 -o gpurun_out/${R}_step_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-refine-run --no-matrix --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/${R}_step_kernels.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_full_step_kernels_raw.csv
ncu -i gpurun_out/${R}_step_kernels.ncu-rep --page details > gpurun_out/${R}_ncu_full_step_kernels_details.txt
python tools/ncu_brief.py gpurun_out/${R}_ncu_full_step_kernels_raw.csv > gpurun_out/${R}_ncu_brief.txt
rm -f gpurun_out/${R}_step_kernels.ncu-rep
cat gpurun_out/${R}_ncu_brief.txt | head -30
# 4. the previous generation is still green
XXM_SCAN_GEN=6 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${R}_pytest_gpu_gen6.log 2>&1; echo "rc $?" >> gpurun_out/${R}_pytest_gpu_gen6.log
tail -2 gpurun_out/${R}_pytest_gpu_gen6.log
# 5. the full bench line and the reference arm
timeout 600 python bench.py > gpurun_out/${R}_bench_ours_1gpu.json 2> gpurun_out/${R}_bench_ours_1gpu.err
timeout 200 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
ls -la gpurun_out | head -40
