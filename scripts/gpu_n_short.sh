#!/bin/bash
# Short N-GPU confirmation of the headline bench (product arm, then the stock arm if time allows).
N=${1:-8}
mkdir -p gpurun_out
for impl in b200 nccl; do
  echo "== bench $impl N=$N"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 4 --warmup 3 --impl $impl --no-e2e > gpurun_out/bench_${impl}_$N.json 2> gpurun_out/bench_${impl}_$N.err
  echo "rc=$?"; tail -2 gpurun_out/bench_${impl}_$N.err | cut -c1-300; tail -1 gpurun_out/bench_${impl}_$N.json | cut -c1-600
done
