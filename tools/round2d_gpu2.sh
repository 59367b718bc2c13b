# 2-GPU sanity pass for the generation-7 scan kernels: the tests that need two devices, the N-rank parity tool, a short 2-rank bench.
R=r02d
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_bound_binary.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/${R}_pytest_2gpu.log 2>&1; echo "rc $?" >> gpurun_out/${R}_pytest_2gpu.log
tail -3 gpurun_out/${R}_pytest_2gpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/comm_parity.py > gpurun_out/${R}_comm_parity_2gpu.json 2> gpurun_out/${R}_comm_parity_2gpu.err; tail -c 600 gpurun_out/${R}_comm_parity_2gpu.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 2 > gpurun_out/${R}_bench_ours_2gpu.json 2> gpurun_out/${R}_bench_ours_2gpu.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/${R}_bench_ours_2gpu.json"))
    print("2 GPUs: ms/step", d["ms_per_step"], "value %.3e" % d["value"], "parity", d.get("parity_nranks"), "strong", d.get("strong_1M", {}).get("ms_per_step"), "e2e ms", d["e2e"]["ms_per_step"])
except Exception as e:
    print("no bench line", e)
P
