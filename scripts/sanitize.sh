#!/bin/bash
# Race / memory checks of the hand-written kernels (absent in the reference).
# Runs a small-shape subset of the single-GPU kernel tests under compute-sanitizer; pass the tools to run
# (default: memcheck racecheck synccheck).  One log + one summary line per tool under gpurun_out/ (copy to profiles/).
TOOLS=${@:-memcheck racecheck synccheck}
mkdir -p gpurun_out
export PYTHONPATH=.
SEL='(sgd or adamw or rope or small_allreduce or (tcgen05_gemm and 128) or (layernorm and 7))'
SEL_CONV='(fprop and 2-64-64-8-8) or (dgrad and 2-64-64-8-8) or (wgrad and 2-64-64-8-8) or (stride2 and 2-64-64-16) or (stem and 2-32) or (statistics and 2-64-64-8-8-1)'
for TOOL in $TOOLS; do
  timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bn.py -m gpu -q -x -k "$SEL" --timeout 500 > gpurun_out/sanitizer_$TOOL.log 2>&1
  rc1=$?
  timeout 600 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 120 \
    python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "$SEL_CONV" --timeout 500 > gpurun_out/sanitizer_${TOOL}_conv.log 2>&1
  rc2=$?
  echo "sanitizer($TOOL) kernels rc=$rc1 conv rc=$rc2 | $(grep -hE 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_$TOOL.log gpurun_out/sanitizer_${TOOL}_conv.log | tr '\n' ';' | cut -c1-300)"
done
