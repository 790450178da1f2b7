#!/bin/bash
# Last validation of round 1 (GPU budget nearly spent): acceptance first, then the Llama / BERT round times.
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== pytest -m gpu"; timeout 150 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-400
echo "== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
rm -f gpurun_out/models_1d.jsonl
echo "== llama"; timeout 100 python scripts/bench_models.py --model llama3_8b_lora --impl b200 --rounds 4 --warmup 3 --out gpurun_out/models_1d.jsonl 2> gpurun_out/model_llama_b200.err | cut -c1-400; tail -1 gpurun_out/model_llama_b200.err | cut -c1-200
echo "== bert"; timeout 60 python scripts/bench_models.py --model bert_base --impl b200 --rounds 5 --warmup 3 --out gpurun_out/models_1d.jsonl 2>/dev/null | cut -c1-300
echo "== bench"; timeout 90 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_last.json
