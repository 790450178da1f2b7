#!/bin/bash
# Round 2, second multi-GPU call (N = $1): data-plane suite again (first run stopped at a tolerance), conv tests with the masked
# residual add, 3-task demo through the resident workers, ResNet-50 bench
N=${1:-2}
TAG=${2:-r2b}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== [0] conv / bn tests"; timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bn.py tests/test_gpu_multi.py -q -m gpu --timeout 300 2>&1 | tail -4 | cut -c1-300
echo "== [1] data plane suite, $N GPUs"
timeout 420 $TR --master-port 29511 tests/dist_comm_check.py --out gpurun_out/comm_${N}gpu_${TAG}.json > gpurun_out/comm_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/comm_${N}gpu_${TAG}.log | cut -c1-2500
echo "== [3] bench.py ResNet-50"
timeout 420 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_resnet50_${N}gpu_${TAG}.json 2> gpurun_out/bench_resnet50_${N}gpu_${TAG}.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_resnet50_${N}gpu_${TAG}.json
echo "== [5] full stack: vserver + $N x vnode --gpu k, three FedAvg tasks"
timeout 420 python scripts/demo_network_gpu.py --nodes $N --model resnet50 --rounds 4 --repeat 3 --out gpurun_out/demo_network_${N}gpu_${TAG}.jsonl 2>&1 | tail -4 | cut -c1-700
