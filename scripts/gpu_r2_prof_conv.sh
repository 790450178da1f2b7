#!/bin/bash
# ncu --set full capture of the implicit-GEMM convolution kernel on 8 representative ResNet-50 layer shapes (batch 64)
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 8 -f -o gpurun_out/igemm_prof_r2a \
    python scripts/conv_probe.py --child profile > gpurun_out/igemm_prof_r2a.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/igemm_prof_r2a.log | cut -c1-300; ls -la gpurun_out/igemm_prof_r2a.ncu-rep
