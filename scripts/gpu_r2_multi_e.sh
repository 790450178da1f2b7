#!/bin/bash
# Round 2, multi-GPU call e (N = $1): K1 with warp-wide batched flag polling; timeline of a ResNet-50 round (1 GPU)
N=${1:-2}
TAG=${2:-r2e}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== [1] data plane suite"
timeout 420 $TR --master-port 29511 tests/dist_comm_check.py --out gpurun_out/comm_${N}gpu_${TAG}.json > gpurun_out/comm_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/comm_${N}gpu_${TAG}.json"))
    print({k: d[k] for k in ("k2_checks_passed", "k1_push", "k1_bcast_gemm_ms", "nccl_bcast_then_cublas_ms", "plain_tcgen05_gemm_ms") if k in d})
except Exception as e:
    print("no result:", e)
PY
tail -3 gpurun_out/comm_${N}gpu_${TAG}.log | cut -c1-400
echo "== [2] K1 engine check"
timeout 300 $TR --master-port 29541 tests/dist_k1_engine_check.py --out gpurun_out/k1_engine_${N}gpu_${TAG}.json > gpurun_out/k1_engine_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
grep -E '^\{|Error|error|assert' gpurun_out/k1_engine_${N}gpu_${TAG}.log | tail -3 | cut -c1-700
if [ "$N" = "2" ]; then
echo "== [3] timeline of one ResNet-50 round (CUPTI, analysis only)"
timeout 300 python scripts/trace_round.py --model resnet50 --out gpurun_out/trace_round_resnet50_${TAG}.json 2>&1 | tail -60 | cut -c1-200
echo "== [4] gemm tests (plain path after the producer restructure)"
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm2.py -q -m gpu -k "gemm" --timeout 200 2>&1 | tail -2
timeout 200 python scripts/kernel_bench.py --only gemm 2>&1 | grep "^{" | cut -c1-250 | head -3
fi
