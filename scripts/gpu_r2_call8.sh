#!/bin/bash
# Round 2, 1-GPU call 8: statistics epilogue with fp64 L2 reductions (no serial fold), BN element-wise variants, 2x2 maxpool backward
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== [2] ResNet-50 layer table"; timeout 900 python scripts/conv_probe.py --time --groups resnet50 --out gpurun_out/conv_probe_r50_r2f.jsonl 2>&1 | grep -E "^# |crashed" | cut -c1-300; python scripts/conv_layer_table.py gpurun_out/conv_probe_r50_r2f.jsonl > gpurun_out/conv_layers_r50_r2f.md 2>&1; tail -8 gpurun_out/conv_layers_r50_r2f.md
echo "== [3] BN passes per variant"; timeout 600 python scripts/bn_bench.py --out gpurun_out/bn_bench_r2f.jsonl 2>&1 | grep -E "per_step_ms|failed" | cut -c1-300
for cfg in 0 1 2; do
  echo "== [4] bench, BN cfg $cfg"; V6B200_BN_CFG=$cfg timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
done
echo "== [4b] separate statistics"; V6B200_CONV_STATS_MIN_KB=1000 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [5] launch list"; V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_resnet50_r2f.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_resnet50_r2f.csv gpurun_out/launches_resnet50_r2f.txt 2>/dev/null | head -16
