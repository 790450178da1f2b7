#!/bin/bash
# Round 2, 1-GPU call 7: role warps in the highest warp ids (scheduler priority), 2 TMA threads, PDL for igemm
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== [2] conv probe timing"; timeout 700 python scripts/conv_probe.py --time --out gpurun_out/conv_probe_time_r2e.jsonl 2>&1 | grep -E "^# [0-9]|crashed" | cut -c1-300
echo "== [3] bench tc (+ comparator arms)"; timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2e_tc.json 2> gpurun_out/bench_r2e_tc.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_r2e_tc.json; tail -3 gpurun_out/bench_r2e_tc.err
echo "== [3b] statistics fused only for >= 8 k-blocks"; V6B200_CONV_STATS_MIN_KB=8 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [3c] separate statistics everywhere"; V6B200_CONV_STATS_MIN_KB=1000 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [3d] no PDL"; V6B200_PDL=0 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [4] launch list of one tc round"; V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_resnet50_r2e.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_resnet50_r2e.csv gpurun_out/launches_resnet50_r2e.txt 2>/dev/null | head -12
echo "== [5] bert: linear bwd tc vs cublas"
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
V6B200_LINEAR_BWD=cublas timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [6] kernel bench gemm"; timeout 300 python scripts/kernel_bench.py --only gemm 2>&1 | grep "^{" | cut -c1-330
echo "== [7] ncu igemm"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 8 -f -o gpurun_out/igemm_prof_r2e python scripts/conv_probe.py --child profile > gpurun_out/igemm_prof_r2e.log 2>&1; echo "rc=$?"
echo "== [8] ResNet-50 layer table"; timeout 900 python scripts/conv_probe.py --time --groups resnet50 --out gpurun_out/conv_probe_r50_r2e.jsonl 2>&1 | grep -E "^# |crashed" | cut -c1-300; python scripts/conv_layer_table.py gpurun_out/conv_probe_r50_r2e.jsonl > gpurun_out/conv_layers_r50_r2e.md 2>&1; tail -8 gpurun_out/conv_layers_r50_r2e.md
echo "== [9] attention bench (roles flipped)"; timeout 300 python scripts/kernel_bench.py --only attn,glm 2>&1 | grep "^{" | cut -c1-400
