#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo "== pytest fedavg engine"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -k "fedavg_engine or small_allreduce or trainer" > gpurun_out/pytest_gpu_engine.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_engine.log
echo "== comm check N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_comm_check.py --out gpurun_out/comm_$N.json > gpurun_out/comm_$N.log 2>&1; echo "comm rc=$?"; grep -v "^\[rank.*Traceback\|^W09\|OMP_NUM\|^\*\*\*" gpurun_out/comm_$N.log | tail -25
