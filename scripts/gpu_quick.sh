#!/bin/bash
# 1-GPU quick iteration: selected GPU tests, kernel bench subset, headline bench (both arms), launch list.
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -k "${1:-bn or glm or rope or zoo or trainer}" > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_quick.log
echo "== kernel bench (glm)"; timeout 300 python scripts/kernel_bench.py --only glm 2>&1 | tail -2 | cut -c1-300
for impl in b200 nccl; do
  echo "== bench $impl"; timeout 600 python bench.py --steps 6 --warmup 3 --impl $impl > gpurun_out/bench_${impl}_1.json 2> gpurun_out/bench_${impl}_1.err; echo "rc=$?"; tail -2 gpurun_out/bench_${impl}_1.err | cut -c1-300; cut -c1-420 gpurun_out/bench_${impl}_1.json
done
echo "== launch list (graph replay round, NVTX-filtered)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "v6_timed/" -c 4000 --csv \
  --log-file gpurun_out/launches_resnet50_fusedbn.csv python bench.py --steps 1 --warmup 1 --no-graph --no-e2e > gpurun_out/ncu_launch2.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_resnet50_fusedbn.csv", errors="replace")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    name = re.sub(r"<.*", "", r[ki])[:70]
    agg[name] += v; cnt[name] += 1
tot = sum(agg.values())
with open("gpurun_out/launches_resnet50_fusedbn_summary.txt", "w") as f:
    f.write(f"total {tot/1e6:.3f} ms over {sum(cnt.values())} launches (first launches of one federated round, eager, serialized under ncu)\n")
    for k, v in agg.most_common(30):
        f.write(f"{v/1e6:9.3f} ms {100*v/tot:5.1f}% x{cnt[k]:4d}  {k}\n")
print(open("gpurun_out/launches_resnet50_fusedbn_summary.txt").read())
PY
