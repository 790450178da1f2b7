#!/bin/bash
# Round 2, 1-GPU validation + measurement call (most important first; every step has its own timeout)
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest conv"; timeout 600 python -m pytest tests/test_gpu_conv.py -q --timeout 200 2>&1 | tail -12 | cut -c1-300
echo "== [2] conv probe timing (numerics + us vs cuDNN)"; timeout 700 python scripts/conv_probe.py --time --out gpurun_out/conv_probe_time_r2b.jsonl 2>&1 | grep -E "^# [0-9]|crashed" | cut -c1-300
echo "== [3] bench tc (+ comparator arms)"; timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2b_tc.json 2> gpurun_out/bench_r2b_tc.err; echo "rc=$?"; cut -c1-900 gpurun_out/bench_r2b_tc.json; tail -3 gpurun_out/bench_r2b_tc.err
echo "== [4] pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== [5] bench cudnn conv path (no comparators)"; V6B200_CONV=cudnn timeout 200 python bench.py --steps 8 --warmup 3 --baselines '' > gpurun_out/bench_r2b_cudnn.json 2> gpurun_out/bench_r2b_cudnn.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_r2b_cudnn.json
echo "== [6] launch list of one tc round"; V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_resnet50_r2b.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_resnet50_r2b.csv gpurun_out/launches_resnet50_r2b.txt 2>/dev/null | head -34
echo "== [7] gemm bench"; timeout 200 python scripts/kernel_bench.py --only gemm 2>&1 | grep "^{" | cut -c1-330
echo "== [8] ncu igemm v2"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 8 -f -o gpurun_out/igemm_prof_r2b python scripts/conv_probe.py --child profile > gpurun_out/igemm_prof_r2b.log 2>&1; echo "rc=$?"
echo "== [9] configs 3-5 (1 GPU)"
for m in bert_base glm llama3_8b_lora; do
  timeout 400 python bench.py --model $m --steps 6 --warmup 3 > gpurun_out/bench_${m}_1gpu_r2b.json 2> gpurun_out/bench_${m}_1gpu_r2b.err; echo "$m rc=$?"; cut -c1-500 gpurun_out/bench_${m}_1gpu_r2b.json; tail -2 gpurun_out/bench_${m}_1gpu_r2b.err | cut -c1-300
done
