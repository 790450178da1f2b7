#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] conv / linear tests"; timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_linear_bwd.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -8 | cut -c1-260
echo "== [2] linear backward micro-benchmark"; bash scripts/gpu_r2_call20.sh 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['M'], r['K_in'], r['N_out'], 'wgrad', r['dw_wgrad_ms'], 'cublas', r['dw_cublas_ms'])"
echo "== [3] ResNet-50 layer table"; timeout 900 python scripts/conv_probe.py --time --groups resnet50 --out gpurun_out/conv_probe_r50_r2g.jsonl 2>&1 | grep -E "^# |crashed" | cut -c1-300; python scripts/conv_layer_table.py gpurun_out/conv_probe_r50_r2g.jsonl > gpurun_out/conv_layers_r50_r2g.md 2>&1; tail -7 gpurun_out/conv_layers_r50_r2g.md
echo "== [4] benches: new WGRAD plan / legacy plan"
timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
V6B200_WGRAD_PLAN=0 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
V6B200_WGRAD_PLAN=0 timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-300
