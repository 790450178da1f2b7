#!/bin/bash
# Round 2, final single-GPU acceptance + evidence run (what the driver does at round end, plus the per-config benches and launch lists)
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== [1] pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu_r2_final.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu_r2_final.txt | cut -c1-300
echo "== [2] smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== [3] bench.py (default = what the driver runs)"; timeout 400 python bench.py > gpurun_out/bench_b200_1gpu_r2_final.json 2> gpurun_out/bench_b200_1gpu_r2_final.err; echo "rc=$?"; cut -c1-420 gpurun_out/bench_b200_1gpu_r2_final.json
echo "== [3b] reference arm"; timeout 60 python bench.py --impl reference | cut -c1-300
echo "== [4] configs 3-5 with their comparator arm"
for m in bert_base glm llama3_8b_lora; do
  timeout 500 python bench.py --model $m --steps 6 --warmup 3 > gpurun_out/bench_${m}_1gpu_r2_final.json 2> gpurun_out/bench_${m}_1gpu_r2_final.err; echo "$m rc=$?"; cut -c1-330 gpurun_out/bench_${m}_1gpu_r2_final.json
done
echo "== [5] launch lists (one eager round under ncu, per-kernel durations)"
V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_resnet50_r2_final.csv python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_resnet50_r2_final.csv gpurun_out/launches_resnet50_r2_final.txt 2>/dev/null | head -8
V6_PROFILE_RANGE=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 8000 --csv --log-file gpurun_out/launches_bert_base_r2_final.csv python bench.py --model bert_base --steps 1 --warmup 3 --no-graph --no-e2e --baselines '' > /dev/null 2>&1; echo "rc=$?"; python scripts/launch_summary.py gpurun_out/launches_bert_base_r2_final.csv gpurun_out/launches_bert_base_r2_final.txt 2>/dev/null | head -14
echo "== [6] kernel bench"; timeout 400 python scripts/kernel_bench.py --only gemm,attn,norm,optim,glm,rope 2>&1 | grep "^{" | cut -c1-420 > gpurun_out/kernel_bench_r2_final.txt; wc -l gpurun_out/kernel_bench_r2_final.txt
echo "== [7] ncu: convolution kernels"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -c 8 -f -o gpurun_out/igemm_prof_r2_final python scripts/conv_probe.py --child profile > gpurun_out/igemm_prof_r2_final.log 2>&1; echo "rc=$?"
