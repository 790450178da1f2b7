#!/bin/bash
# Race / memory checks of the hand-written kernels (absent in the reference).
# Runs the single-GPU kernel tests under compute-sanitizer; pass a tool: memcheck|racecheck|synccheck|initcheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "sgd or adamw or layernorm and 7 or rope or glm or small_allreduce or tcgen05_gemm and 128" \
  > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "sanitizer($TOOL) rc=$?"; tail -15 gpurun_out/sanitizer_$TOOL.log
