#!/bin/bash
# Round 2, third multi-GPU call (N = $1): K1 on the engine path (bcast="fused") against the plain path, K1 push micro-benchmark
# with the wider push units, BERT-base with and without the fused broadcast
N=${1:-2}
TAG=${2:-r2c}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== [1] K1 engine check, $N GPUs"
timeout 300 $TR --master-port 29541 tests/dist_k1_engine_check.py --out gpurun_out/k1_engine_${N}gpu_${TAG}.json > gpurun_out/k1_engine_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
grep -E '^\{|Error|error|assert' gpurun_out/k1_engine_${N}gpu_${TAG}.log | tail -6 | cut -c1-1500
echo "== [2] data plane suite"
timeout 420 $TR --master-port 29511 tests/dist_comm_check.py --out gpurun_out/comm_${N}gpu_${TAG}.json > gpurun_out/comm_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/comm_${N}gpu_${TAG}.json"))
    print({k: d[k] for k in ("k2_checks_passed", "k1_push", "k1_bcast_gemm_ms", "nccl_bcast_then_cublas_ms", "k2_sharded_mc_ms", "k2_sharded_p2p_ms") if k in d})
except Exception as e:
    print("no result:", e)
PY
echo "== [3] BERT-base: push vs fused"
for b in push fused; do
  timeout 400 $TR --master-port 29515 bench.py --gpus $N --model bert_base --steps 6 --warmup 3 --bcast $b --baselines '' --no-e2e > gpurun_out/bench_bert_base_${N}gpu_${TAG}_$b.json 2> gpurun_out/bench_bert_base_${N}gpu_${TAG}_$b.err; echo "$b rc=$?"; cut -c1-330 gpurun_out/bench_bert_base_${N}gpu_${TAG}_$b.json; tail -2 gpurun_out/bench_bert_base_${N}gpu_${TAG}_$b.err | cut -c1-300
done
