#!/bin/bash
# Round 2, 1-GPU call 6: shared-address-space fix in every kernel + igemm v3.3
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== [2] conv probe timing"; timeout 700 python scripts/conv_probe.py --time --out gpurun_out/conv_probe_time_r2d.jsonl 2>&1 | grep -E "^# [0-9]|crashed" | cut -c1-300
echo "== [3] bench tc (+ comparator arms)"; timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2d_tc.json 2> gpurun_out/bench_r2d_tc.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_r2d_tc.json; tail -3 gpurun_out/bench_r2d_tc.err
echo "== [3b] bench tc, statistics fused only for >= 8 k-blocks"; V6B200_CONV_STATS_MIN_KB=8 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [3c] bench tc, separate statistics everywhere"; V6B200_CONV_STATS_MIN_KB=1000 timeout 300 python bench.py --steps 8 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [4] launch list of one tc round"; V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_resnet50_r2d.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_resnet50_r2d.csv gpurun_out/launches_resnet50_r2d.txt 2>/dev/null | head -12
echo "== [5] bert: linear bwd tc vs cublas"
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
V6B200_LINEAR_BWD=cublas timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [6] kernel bench gemm / attn"; timeout 300 python scripts/kernel_bench.py --only gemm,attn 2>&1 | grep "^{" | cut -c1-420
echo "== [7] attention bwd native vs flash-attn"; V6B200_ATTN_BWD=native timeout 200 python -m pytest tests/test_gpu_attention_bwd.py -q -s --timeout 200 2>&1 | tail -12 | cut -c1-300
echo "== [8] ncu igemm"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 8 -f -o gpurun_out/igemm_prof_r2d python scripts/conv_probe.py --child profile > gpurun_out/igemm_prof_r2d.log 2>&1; echo "rc=$?"
