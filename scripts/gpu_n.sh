#!/bin/bash
# N-GPU run: comm correctness/bandwidth then the headline bench for both arms.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== comm check N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_comm_check.py --out gpurun_out/comm_$N.json > gpurun_out/comm_$N.log 2>&1; echo "comm rc=$?"; grep -E "^\{|Error|error" gpurun_out/comm_$N.log | tail -5
for impl in b200 nccl; do
  echo "== bench $impl N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 6 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_$N.json 2> gpurun_out/bench_${impl}_$N.err
  echo "rc=$?"; tail -2 gpurun_out/bench_${impl}_$N.err | cut -c1-300; tail -1 gpurun_out/bench_${impl}_$N.json | cut -c1-600
done
