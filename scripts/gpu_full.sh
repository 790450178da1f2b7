#!/bin/bash
# 1-GPU acceptance run: what the driver runs at round end (pytest -m gpu, smoke, bench both arms) + the secondary models.
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
for impl in b200 nccl; do
  echo "== bench $impl"; timeout 600 python bench.py --steps 6 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_1.json 2> gpurun_out/bench_${impl}_1.err; echo "rc=$?"; tail -2 gpurun_out/bench_${impl}_1.err | cut -c1-300; cut -c1-330 gpurun_out/bench_${impl}_1.json
done
echo "== bench reference arm"; timeout 300 python bench.py --impl reference | cut -c1-300
rm -f gpurun_out/models_1c.jsonl
for m in bert_base llama3_8b_lora glm; do
  echo "== model $m b200"; timeout 600 python scripts/bench_models.py --model $m --impl b200 --rounds 5 --warmup 3 --out gpurun_out/models_1c.jsonl 2> gpurun_out/model_${m}_b200.err | cut -c1-420; tail -2 gpurun_out/model_${m}_b200.err | cut -c1-300
done
echo "== launch list bert_base (one eager round)"
V6_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 6000 --csv \
  --log-file gpurun_out/launches_bert2.csv python scripts/bench_models.py --model bert_base --impl b200 --rounds 1 --warmup 2 --no-graph > gpurun_out/ncu_launch_bert2.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_bert2.csv gpurun_out/launches_bert2_summary.txt | head -24
