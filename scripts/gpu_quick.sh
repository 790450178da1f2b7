#!/bin/bash
# 1-GPU quick iteration: selected GPU tests, headline bench (both arms), launch list of one full eager round
# (forward + backward + optimizer + aggregation; cudaProfilerStart/Stop range so the autograd thread is included).
mkdir -p gpurun_out
TAG=${2:-v3}
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -k "${1:-bn or resnet or zoo or trainer or optim}" > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_quick.log
for impl in b200 nccl; do
  echo "== bench $impl"; timeout 600 python bench.py --steps 6 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_1.json 2> gpurun_out/bench_${impl}_1.err; echo "rc=$?"; tail -2 gpurun_out/bench_${impl}_1.err | cut -c1-300; cut -c1-420 gpurun_out/bench_${impl}_1.json
done
echo "== launch list (one eager round, profiler range)"
V6_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv \
  --log-file gpurun_out/launches_resnet50_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_resnet50_${TAG}.csv gpurun_out/launches_resnet50_${TAG}_summary.txt
