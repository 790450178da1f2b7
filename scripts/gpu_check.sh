#!/bin/bash
# Single-GPU bring-up on the B200 box: smoke, GPU tests, 1-GPU bench. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench b200"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b200_1.json 2> gpurun_out/bench_b200_1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b200_1.err; cat gpurun_out/bench_b200_1.json
echo "== bench nccl baseline"; timeout 900 python bench.py --impl nccl --steps 5 --warmup 3 > gpurun_out/bench_nccl_1.json 2> gpurun_out/bench_nccl_1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_nccl_1.err; cat gpurun_out/bench_nccl_1.json
