#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] tests"; timeout 600 python -m pytest tests/test_gpu_linear_bwd.py tests/test_gpu_kernels.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -25 | cut -c1-220
echo "== [2] BERT-base (+ comparator arm)"; timeout 400 python bench.py --model bert_base --steps 6 --warmup 3 > gpurun_out/bench_bert_base_1gpu_r2_final2.json 2>/dev/null; cut -c1-330 gpurun_out/bench_bert_base_1gpu_r2_final2.json
echo "== [3] launch list"; V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_bert_base_r2_final2.csv python bench.py --model bert_base --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_bert_base_r2_final2.csv gpurun_out/launches_bert_base_r2_final2.txt 2>/dev/null | head -16
