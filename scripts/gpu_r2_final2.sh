#!/bin/bash
# Round 2, final multi-GPU sanity (N = $1) on the final tree
N=${1:-2}
TAG=r2_final
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== [0] multi-GPU pytest (collects the torchrun checks)"; timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 500 2>&1 | tail -3 | cut -c1-300
echo "== [1] bench.py ResNet-50"
timeout 420 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_resnet50_${N}gpu_${TAG}.json 2> gpurun_out/bench_resnet50_${N}gpu_${TAG}.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_resnet50_${N}gpu_${TAG}.json
echo "== [2] K1 engine check"
timeout 300 $TR --master-port 29541 tests/dist_k1_engine_check.py --out gpurun_out/k1_engine_${N}gpu_${TAG}.json > gpurun_out/k1_engine_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
grep -E '^\{|Error|error|assert' gpurun_out/k1_engine_${N}gpu_${TAG}.log | tail -3 | cut -c1-600
echo "== [3] BERT-base, GLM"
timeout 400 $TR --master-port 29515 bench.py --gpus $N --model bert_base --steps 6 --warmup 3 > gpurun_out/bench_bert_base_${N}gpu_${TAG}.json 2> gpurun_out/bench_bert_base_${N}gpu_${TAG}.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_bert_base_${N}gpu_${TAG}.json
timeout 200 $TR --master-port 29519 bench.py --gpus $N --model glm --steps 6 --warmup 3 > gpurun_out/bench_glm_${N}gpu_${TAG}.json 2> gpurun_out/bench_glm_${N}gpu_${TAG}.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_glm_${N}gpu_${TAG}.json
echo "== [4] full stack"
timeout 420 python scripts/demo_network_gpu.py --nodes $N --model resnet50 --rounds 4 --repeat 2 --out gpurun_out/demo_network_${N}gpu_${TAG}.jsonl 2>&1 | tail -2 | cut -c1-500
