#!/bin/bash
# Round 2 multi-GPU call: N = $1 GPUs of one box.  Order = most important first (the call may be cut short).
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== [1] data plane suite (K2 / K3 / K1 pull + push), $N GPUs"
timeout 420 $TR --master-port 29511 tests/dist_comm_check.py --out gpurun_out/comm_${N}gpu_${TAG}.json > gpurun_out/comm_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/comm_${N}gpu_${TAG}.log | cut -c1-1500
echo "== [2] kill a rank mid-round (sharded / weights, central / delta)"
timeout 120 $TR --master-port 29533 tests/dist_fault_check.py --out gpurun_out/fault_${N}gpu_sharded --server-mode sharded --upload weights_f32 2>&1 | grep -E '^\{' | cut -c1-500
timeout 120 $TR --master-port 29535 tests/dist_fault_check.py --out gpurun_out/fault_${N}gpu_central --server-mode central --upload delta_f32 2>&1 | grep -E '^\{' | cut -c1-500
echo "== [3] bench.py ResNet-50, product + comparator arms"
timeout 420 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_resnet50_${N}gpu_${TAG}.json 2> gpurun_out/bench_resnet50_${N}gpu_${TAG}.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_resnet50_${N}gpu_${TAG}.json
echo "== [4] configs 3-5"
for m in bert_base glm llama3_8b_lora; do
  timeout 400 $TR --master-port 29515 bench.py --gpus $N --model $m --steps 6 --warmup 3 > gpurun_out/bench_${m}_${N}gpu_${TAG}.json 2> gpurun_out/bench_${m}_${N}gpu_${TAG}.err; echo "$m rc=$?"; cut -c1-600 gpurun_out/bench_${m}_${N}gpu_${TAG}.json
done
echo "== [5] full stack: vserver + $N x vnode --gpu k, three FedAvg tasks (the later ones reuse the resident GPU workers)"
timeout 420 python scripts/demo_network_gpu.py --nodes $N --model resnet50 --rounds 4 --repeat 3 --out gpurun_out/demo_network_${N}gpu_${TAG}.jsonl 2>&1 | tail -4 | cut -c1-600
