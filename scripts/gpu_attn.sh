#!/bin/bash
# 1-GPU: attention-focused check (both forward designs), then the transformer model benches.
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== pytest attention"; timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_attention_bwd.py -m gpu -q --timeout 300 > gpurun_out/pytest_attn.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_attn.log | cut -c1-300
echo "== kernel bench attn"; timeout 600 python scripts/kernel_bench.py --only attn > gpurun_out/kernel_bench_attn.log 2>&1; echo "rc=$?"
python - <<'PY'
import ast
for line in open("gpurun_out/kernel_bench_attn.log"):
    if line.startswith("{'B'"):
        d = ast.literal_eval(line)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ("B", "S", "Hq", "D", "causal", "fwd_ms", "fwd_tflops", "fwd2_ms", "fwd2_tflops", "flash_attn2_ms", "sdpa_ms", "fwd2_error")})
PY
cp gpurun_out/kernel_bench.json gpurun_out/kernel_bench_attn.json 2>/dev/null
rm -f gpurun_out/models_1b.jsonl
for m in bert_base llama3_8b_lora; do
  echo "== model $m b200"; timeout 600 python scripts/bench_models.py --model $m --impl b200 --rounds 5 --warmup 3 --out gpurun_out/models_1b.jsonl 2> gpurun_out/model_${m}_b200.err | cut -c1-400; tail -2 gpurun_out/model_${m}_b200.err | cut -c1-300
done
echo "== pytest bn/resnet (PDL launches)"; timeout 900 python -m pytest tests/test_gpu_bn.py tests/test_gpu_resnet_ops.py -m gpu -q --timeout 300 > gpurun_out/pytest_bn.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_bn.log | cut -c1-300
for pdl in 1 0; do
  echo "== bench b200 PDL=$pdl"; V6B200_PDL=$pdl timeout 600 python bench.py --steps 6 --warmup 3 --impl b200 > gpurun_out/bench_b200_pdl$pdl.json 2> gpurun_out/bench_b200_pdl$pdl.err; echo "rc=$?"; tail -2 gpurun_out/bench_b200_pdl$pdl.err | cut -c1-300; cut -c1-330 gpurun_out/bench_b200_pdl$pdl.json
done
echo "== launch list bert_base (one eager round)"
V6_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
  --log-file gpurun_out/launches_bert.csv python scripts/bench_models.py --model bert_base --impl b200 --rounds 1 --warmup 2 --no-graph > gpurun_out/ncu_launch_bert.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_bert.csv gpurun_out/launches_bert_summary.txt | head -32
echo "== glm test + bench"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "glm" 2>&1 | tail -2; timeout 300 python scripts/kernel_bench.py --only glm 2>&1 | grep "rows" | cut -c1-300
