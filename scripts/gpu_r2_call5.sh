#!/bin/bash
# Round 2, 1-GPU call 5: igemm v3 (EPI templates, paired chunks, block_n cost model)
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest conv"; timeout 600 python -m pytest tests/test_gpu_conv.py -q --timeout 200 2>&1 | tail -6 | cut -c1-300
echo "== [2] conv probe timing"; timeout 700 python scripts/conv_probe.py --time --out gpurun_out/conv_probe_time_r2c.jsonl 2>&1 | grep -E "^# [0-9]|crashed" | cut -c1-300
echo "== [3] bench tc (+ comparator arms)"; timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2c_tc.json 2> gpurun_out/bench_r2c_tc.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_r2c_tc.json; tail -3 gpurun_out/bench_r2c_tc.err
echo "== [4] launch list of one tc round"; V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_resnet50_r2c.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_resnet50_r2c.csv gpurun_out/launches_resnet50_r2c.txt 2>/dev/null | head -16
echo "== [5] bert: linear bwd tc vs cublas"
timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
V6B200_LINEAR_BWD=cublas timeout 300 python bench.py --model bert_base --steps 6 --warmup 3 --baselines '' --no-e2e 2>/dev/null | cut -c1-330
echo "== [6] ncu igemm v3"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 8 -f -o gpurun_out/igemm_prof_r2c python scripts/conv_probe.py --child profile > gpurun_out/igemm_prof_r2c.log 2>&1; echo "rc=$?"
echo "== [7] pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== [8] sanitizer"; bash scripts/sanitize.sh memcheck racecheck 2>&1 | tail -4 | cut -c1-400
