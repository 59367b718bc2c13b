# Round profile recipe (run on the GPU box: gpurun -- bash tools/profile_round.sh [rNN]): bench line, reference arm, launch
# lists, one full ncu capture of the kernels of a refinement step, the DRAM traffic of a scan, and the memcheck / racecheck
# runs.  Each ncu pass runs after the same command has exited 0 without ncu.
R=${1:-r02}
set -x
timeout 900 python bench.py > gpurun_out/${R}_bench_ours_1gpu.json 2> gpurun_out/${R}_bench_ours_1gpu.err || exit 1
timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
# every launch of a short bench command (steps, one e2e load + step, the whole refinement run)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-matrix --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
# one launch of each kernel of a step, full set
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_scan5|k_tile_scan|k_expand7|k_select_cls_fwd|k_count_pwm3" -s 10 -c 5 -o gpurun_out/${R}_step_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-refine-run --no-matrix --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/${R}_step_kernels.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_full_step_kernels_raw.csv
ncu -i gpurun_out/${R}_step_kernels.ncu-rep --page details > gpurun_out/${R}_ncu_full_step_kernels_details.txt
python tools/ncu_brief.py gpurun_out/${R}_ncu_full_step_kernels_raw.csv > gpurun_out/${R}_ncu_brief.txt
# the --revcomp select kernel and the kernels of the other workloads: launch lists of two steps each
for w in "cfg2 1000000" "cfg5 2000000" "cols12 1000000" "cols17 1000000"; do
  set -- $w
  python tools/step_workload.py $1 $2 5 2>&1 | tail -1 > gpurun_out/${R}_step_$1.txt
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${R}_launches_$1_steps.csv python tools/step_workload.py $1 $2 2 > /dev/null 2>&1
done
# memcheck + racecheck over the smoke run and the ragged / low-complexity stress case
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_memcheck_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/${R}_sanitizer_memcheck_smoke.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -q -m gpu -k "low_complexity_ragged" > gpurun_out/${R}_sanitizer_memcheck_stress.txt 2>&1; echo "exit $?" >> gpurun_out/${R}_sanitizer_memcheck_stress.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_racecheck_smoke.txt 2>&1; echo "exit $?" >> gpurun_out/${R}_sanitizer_racecheck_smoke.txt
tail -3 gpurun_out/${R}_sanitizer_*.txt
