#!/bin/bash
# Minimal acceptance: what the driver runs at round end (pytest -m gpu, smoke, 1-GPU bench of the product arm).
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
echo "== bench"; timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_default.json
echo "== glm"; timeout 200 python scripts/bench_models.py --model glm --impl b200 --rounds 5 --warmup 3 2>/dev/null | cut -c1-300
