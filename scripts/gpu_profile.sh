#!/bin/bash
# 1-GPU: full GPU test suite, per-kernel roofline bench, launch list of one ResNet-50 round, ncu --set full
# captures of the top hand-written kernels. Everything lands in gpurun_out/.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== kernel bench"; timeout 900 python scripts/kernel_bench.py > gpurun_out/kernel_bench.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/kernel_bench.log | cut -c1-400
echo "== launch list (eager round, NVTX-filtered)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "v6_timed/" -c 1500 --csv \
  --log-file gpurun_out/launches_resnet50_eager.csv python bench.py --steps 1 --warmup 1 --no-graph --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_resnet50_eager.csv", errors="replace")) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    name = re.sub(r"<.*", "", r[ki])[:70]
    agg[name] += v; cnt[name] += 1
tot = sum(agg.values())
with open("gpurun_out/launches_resnet50_summary.txt", "w") as f:
    f.write(f"total {tot/1e6:.3f} ms over {sum(cnt.values())} launches (one federated round, eager, serialized under ncu)\n")
    for k, v in agg.most_common(25):
        f.write(f"{v/1e6:9.3f} ms {100*v/tot:5.1f}% x{cnt[k]:4d}  {k}\n")
print(open("gpurun_out/launches_resnet50_summary.txt").read())
PY
echo "== ncu full: tcgen05 gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 2 -c 1 -o gpurun_out/prof_gemm -f \
  python -c "
import torch; from vantage6_b200.ops import gemm as G
a=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16); w=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16)
for _ in range(4): G.gemm_bf16(a,w)
torch.cuda.synchronize()" > gpurun_out/ncu_gemm.log 2>&1; echo "rc=$?"
echo "== ncu full: fedavg_round (world 1) + flat sgd + layernorm + glm + flash attn"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fedavg_round_kernel|flat_optim_kernel|norm_fwd_kernel|norm_bwd_kernel|glm_logistic_kernel|flash_fwd_kernel|rope_kernel" -c 12 -o gpurun_out/prof_misc -f \
  python -c "
import torch
from vantage6_b200.parallel.fedavg import FedAvgEngine
from vantage6_b200.ops import optim as O, norm as N, glm as K8, attention as A, rope as R
dev=torch.device('cuda',0)
e=FedAvgEngine(25610152,0,1,dev,data_plane='native'); e.w.normal_(); e.initialize_global(); e.aggregate(1.0)
w=torch.randn(25557032,device=dev); g=torch.randn_like(w); O.FlatSGD(w).step(g)
x=torch.randn(32768,768,device=dev,dtype=torch.bfloat16,requires_grad=True); gm=torch.ones(768,device=dev,requires_grad=True); b=torch.zeros(768,device=dev,requires_grad=True)
y,_=N.layer_norm(x,gm,b); y.sum().backward()
X=torch.randn(1000000,256,device=dev,dtype=torch.bfloat16); yy=(torch.rand(1000000,device=dev)<0.5).float(); K8.logistic_grad(X,yy,torch.zeros(257,device=dev))
q=torch.randn(4,2048,32,128,device=dev,dtype=torch.bfloat16); k=torch.randn(4,2048,8,128,device=dev,dtype=torch.bfloat16); v=torch.randn_like(k)
A.flash_attn_fwd(q,k,v,True)
cos,sin=R.rope_tables(2048,128,device=dev); R.apply_rope(q,k,cos,sin)
torch.cuda.synchronize()" > gpurun_out/ncu_misc.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_misc.log
echo "== batch sweep"
for bs in 128 256; do for impl in b200 nccl; do
  timeout 600 python bench.py --steps 4 --warmup 3 --impl $impl --batch $bs --local-samples $((bs*8)) --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$impl', $bs, round(d['images_per_sec']), round(d['ms_per_step'],1))"
done; done | tee gpurun_out/batch_sweep.txt
