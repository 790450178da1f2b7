#!/bin/bash
# GLM iteration as two launches (gradient kernel + fold / NVLink all-reduce / update kernel): validation + A/B at N = $1 GPUs
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
echo "== [1] GLM + attention tests incl. opt-in"; V6B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_glm.py tests/test_gpu_attention.py -q -m gpu --timeout 300 2>&1 | grep -v "^E   \s*+" | tail -8 | cut -c1-220
fi
echo "== [2] GLM bench, fused iteration"; V6B200_GLM_FUSED=1 timeout 200 $TR --master-port 29519 bench.py --gpus $N --model glm --steps 6 --warmup 3 > gpurun_out/bench_glm_${N}gpu_r2_fused.json 2> gpurun_out/bench_glm_${N}gpu_r2_fused.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_glm_${N}gpu_r2_fused.json; tail -2 gpurun_out/bench_glm_${N}gpu_r2_fused.err | cut -c1-300
echo "== [3] GLM bench, composed path"; V6B200_GLM_FUSED=0 timeout 200 $TR --master-port 29521 bench.py --gpus $N --model glm --steps 6 --warmup 3 --baselines '' > gpurun_out/bench_glm_${N}gpu_r2_plain.json 2>/dev/null; cut -c1-330 gpurun_out/bench_glm_${N}gpu_r2_plain.json
if [ "$N" = "8" ]; then
echo "== [4] ResNet-50 at 8 GPUs, final tree"; timeout 300 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_resnet50_8gpu_r2_final.json 2> gpurun_out/bench_resnet50_8gpu_r2_final.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_resnet50_8gpu_r2_final.json
echo "== [5] BERT-base at 8 GPUs, final tree"; timeout 300 $TR --master-port 29515 bench.py --gpus 8 --model bert_base --steps 6 --warmup 3 > gpurun_out/bench_bert_base_8gpu_r2_final.json 2> gpurun_out/bench_bert_base_8gpu_r2_final.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_bert_base_8gpu_r2_final.json
fi
