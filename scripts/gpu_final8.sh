#!/bin/bash
# 8-GPU (or N-GPU) final run: comm check, headline bench both arms, secondary models, full-stack demo network.
N=${1:-8}
mkdir -p gpurun_out
export PYTHONPATH=.
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== comm check N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tests/dist_comm_check.py --out gpurun_out/comm_$N.json > gpurun_out/comm_$N.log 2>&1; echo "comm rc=$?"; grep -E "^\{|Error|error|FAIL|passed" gpurun_out/comm_$N.log | tail -6 | cut -c1-600
for impl in b200 nccl; do
  echo "== bench $impl N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 6 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_$N.json 2> gpurun_out/bench_${impl}_$N.err
  echo "rc=$?"; tail -2 gpurun_out/bench_${impl}_$N.err | cut -c1-300; tail -1 gpurun_out/bench_${impl}_$N.json | cut -c1-700
done
rm -f gpurun_out/models_$N.jsonl
for m in bert_base glm; do for impl in b200 nccl; do
  echo "== model $m $impl N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    scripts/bench_models.py --model $m --impl $impl --rounds 5 --warmup 3 --out gpurun_out/models_$N.jsonl 2> gpurun_out/model_${m}_${impl}_$N.err | cut -c1-500
  tail -1 gpurun_out/model_${m}_${impl}_$N.err | cut -c1-200
done; done
echo "== full stack: vserver + $N x vnode --gpu k, FedAvg(resnet50) + GLM through the control plane"
rm -f gpurun_out/demo_network_$N.jsonl; timeout 420 python scripts/demo_network_gpu.py --nodes $N --model resnet50 --rounds 4 --glm --out gpurun_out/demo_network_$N.jsonl 2>&1 | tail -6 | cut -c1-900
