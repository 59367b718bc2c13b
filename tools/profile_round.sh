# Round profile recipe (run on the GPU box: gpurun -- bash tools/profile_round.sh [rNN]): bench line, reference arm, launch
# list, one full ncu capture of the kernels of a refinement step, and the exports that go to profiles/.  Each ncu pass runs
# after the same command has exited 0 without ncu.
R=${1:-r01}
set -x
timeout 400 python bench.py > gpurun_out/${R}_bench_ours_1gpu.json 2> gpurun_out/${R}_bench_ours_1gpu.err || exit 1
timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
# every launch of a short bench command (steps, one e2e load + step, the whole refinement run)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
# one launch of each kernel of a step, full set
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_scan5|k_tile_scan|k_expand6|k_select_cls_fwd|k_count_pwm3" -s 10 -c 5 -o gpurun_out/${R}_step_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-refine-run --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/${R}_step_kernels.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_full_step_kernels_raw.csv
ncu -i gpurun_out/${R}_step_kernels.ncu-rep --page details > gpurun_out/${R}_ncu_full_step_kernels_details.txt
python tools/ncu_brief.py gpurun_out/${R}_ncu_full_step_kernels_raw.csv > gpurun_out/${R}_ncu_brief.txt
# host side of the e2e number
timeout 200 python tools/load_probe.py > gpurun_out/${R}_load_probe.txt 2>&1
