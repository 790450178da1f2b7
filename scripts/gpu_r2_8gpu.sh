#!/bin/bash
# Round 2, the 8-GPU call.  Most important first (the call may be cut short by the budget).
N=8
TAG=${1:-r2}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_8_${TAG}.txt 2>&1
echo "== [1] bench.py ResNet-50, product + comparator arms"
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_resnet50_${N}gpu_${TAG}.json 2> gpurun_out/bench_resnet50_${N}gpu_${TAG}.err; echo "rc=$?"; cut -c1-420 gpurun_out/bench_resnet50_${N}gpu_${TAG}.json
echo "== [2] data plane suite (K2 / K3 / K1 pull + push)"
timeout 300 $TR --master-port 29511 tests/dist_comm_check.py --out gpurun_out/comm_${N}gpu_${TAG}.json > gpurun_out/comm_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/comm_${N}gpu_${TAG}.log | cut -c1-2600
echo "== [3] K1 on the engine path"
timeout 200 $TR --master-port 29541 tests/dist_k1_engine_check.py --out gpurun_out/k1_engine_${N}gpu_${TAG}.json > gpurun_out/k1_engine_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
grep -E '^\{|Error|error|assert' gpurun_out/k1_engine_${N}gpu_${TAG}.log | tail -4 | cut -c1-1500
echo "== [4] kill a rank mid-round"
timeout 120 $TR --master-port 29533 tests/dist_fault_check.py --out gpurun_out/fault_${N}gpu_sharded --server-mode sharded --upload weights_f32 2>&1 | grep -E '^\{' | cut -c1-600
echo "== [5] configs 3-5"
timeout 300 $TR --master-port 29515 bench.py --gpus $N --model bert_base --steps 6 --warmup 3 > gpurun_out/bench_bert_base_${N}gpu_${TAG}.json 2> gpurun_out/bench_bert_base_${N}gpu_${TAG}.err; echo "bert rc=$?"; cut -c1-420 gpurun_out/bench_bert_base_${N}gpu_${TAG}.json
timeout 300 $TR --master-port 29517 bench.py --gpus $N --model bert_base --steps 6 --warmup 3 --bcast fused --baselines '' --no-e2e > gpurun_out/bench_bert_base_${N}gpu_${TAG}_fused.json 2> gpurun_out/bench_bert_base_${N}gpu_${TAG}_fused.err; echo "bert fused rc=$?"; cut -c1-330 gpurun_out/bench_bert_base_${N}gpu_${TAG}_fused.json
timeout 200 $TR --master-port 29519 bench.py --gpus $N --model glm --steps 6 --warmup 3 > gpurun_out/bench_glm_${N}gpu_${TAG}.json 2> gpurun_out/bench_glm_${N}gpu_${TAG}.err; echo "glm rc=$?"; cut -c1-420 gpurun_out/bench_glm_${N}gpu_${TAG}.json
echo "== [6] full stack: vserver + 8 x vnode --gpu k, three FedAvg tasks"
timeout 300 python scripts/demo_network_gpu.py --nodes $N --model resnet50 --rounds 4 --repeat 3 --out gpurun_out/demo_network_${N}gpu_${TAG}.jsonl 2>&1 | tail -3 | cut -c1-700
echo "== [7] Llama-3 8B LoRA"
timeout 400 $TR --master-port 29521 bench.py --gpus $N --model llama3_8b_lora --steps 4 --warmup 3 > gpurun_out/bench_llama3_8b_lora_${N}gpu_${TAG}.json 2> gpurun_out/bench_llama3_8b_lora_${N}gpu_${TAG}.err; echo "llama rc=$?"; cut -c1-420 gpurun_out/bench_llama3_8b_lora_${N}gpu_${TAG}.json
