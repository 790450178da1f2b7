#!/bin/bash
# Round 2, fourth multi-GPU call (N = $1): K1 engine check (fixed driver), K1 push with 5 pusher warps per CTA
N=${1:-2}
TAG=${2:-r2d}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== [1] K1 engine check, $N GPUs"
timeout 300 $TR --master-port 29541 tests/dist_k1_engine_check.py --out gpurun_out/k1_engine_${N}gpu_${TAG}.json > gpurun_out/k1_engine_${N}gpu_${TAG}.log 2>&1; echo "rc=$?"
grep -E '^\{|Error|error|assert' gpurun_out/k1_engine_${N}gpu_${TAG}.log | tail -6 | cut -c1-2000
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
echo "== [3] 1-GPU: conv / bn tests + ResNet bench (stem statistics fused, PDL for bn_apply / bn_bwd_reduce)"
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_bn.py tests/test_gpu_resnet_ops.py tests/test_gpu_gemm2.py tests/test_gpu_kernels.py -q -m gpu --timeout 300 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
V6B200_PDL=0 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
