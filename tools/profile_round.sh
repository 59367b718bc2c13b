set -x
timeout 300 python bench.py > gpurun_out/r02_bench_ours_1gpu.json 2> gpurun_out/r02_bench_ours_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_scan5|k_expand5|k_select_mops_fwd|k_count_pwm2" -s 12 -c 4 -o gpurun_out/r02_scan_step python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
