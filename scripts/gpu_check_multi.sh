#!/bin/bash
# Multi-GPU bring-up: N = number of GPUs (arg 1). Comm correctness + bandwidth, then bench at N.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== pytest gpu (no -x)"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== comm check N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_comm_check.py --out gpurun_out/comm_$N.json > gpurun_out/comm_$N.log 2>&1; echo "comm rc=$?"; tail -30 gpurun_out/comm_$N.log
for impl in b200 nccl; do
  echo "== bench $impl N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 5 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_$N.json 2> gpurun_out/bench_${impl}_$N.err
  echo "rc=$?"; tail -3 gpurun_out/bench_${impl}_$N.err; tail -1 gpurun_out/bench_${impl}_$N.json
done
