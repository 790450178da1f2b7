#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== conv probe numerics (v2 kernel)"; timeout 600 python scripts/conv_probe.py --out gpurun_out/conv_probe_r2b.jsonl 2>&1 | grep -E "^#|false|crashed" | cut -c1-400
echo "== conv probe timing"; timeout 600 python scripts/conv_probe.py --time --out gpurun_out/conv_probe_time_r2b.jsonl 2>&1 | grep -E "^# [0-9]|crashed" | cut -c1-300
echo "== pytest conv"; timeout 600 python -m pytest tests/test_gpu_conv.py -x -q --timeout 200 2>&1 | tail -15 | cut -c1-300
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench tc"; timeout 300 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r2b_tc.json 2> gpurun_out/bench_r2b_tc.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_r2b_tc.json; tail -3 gpurun_out/bench_r2b_tc.err
echo "== bench cudnn"; V6B200_CONV=cudnn timeout 300 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_r2b_cudnn.json 2> gpurun_out/bench_r2b_cudnn.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_r2b_cudnn.json
