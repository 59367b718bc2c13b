# Round profile recipe (run on the GPU box: gpurun -- bash tools/profile_round.sh): bench line, launch list, one full ncu
# capture of the kernels of a refinement step.  Each ncu pass runs after the same command has exited 0 without ncu.
set -x
timeout 400 python bench.py > gpurun_out/r01_bench_ours_1gpu.json 2> gpurun_out/r01_bench_ours_1gpu.err || exit 1
timeout 300 python bench.py --impl reference > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_scan5|k_tile_scan|k_expand6|k_select_cls_fwd|k_count_pwm3" -s 10 -c 5 -o gpurun_out/r01_step_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
