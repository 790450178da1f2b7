#!/bin/bash
# 1-GPU: ncu --set full captures of the kernels added late in round 1: two-CTA-per-SM attention forward, native
# attention backward, fused bias/GELU backward, K3 small all-reduce (world 1), K8 GLM v4.
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flash_fwd2_kernel|flash_bwd_dq_kernel|flash_bwd_dkv_kernel|bias_act_bwd_kernel|small_allreduce_kernel|glm_logistic_kernel|norm_param_grad_fold" -c 8 -o gpurun_out/prof_late -f \
  python -c "
import torch
from vantage6_b200.ops import attention as A, gemm as G, glm as K8
from vantage6_b200.parallel.fedavg import SmallAggregator
dev=torch.device('cuda',0)
q=torch.randn(4,2048,32,128,device=dev,dtype=torch.bfloat16); k=torch.randn(4,2048,8,128,device=dev,dtype=torch.bfloat16); v=torch.randn_like(k)
o,lse=A.flash_attn_fwd(q,k,v,True,variant='2cta')
A.flash_attn_bwd(torch.randn_like(o),q,k,v,o,lse,True)
dy=torch.randn(4096,3072,device=dev).to(torch.bfloat16); pre=torch.randn(4096,3072,device=dev).to(torch.bfloat16)
b=torch.nn.Parameter(torch.zeros(3072,device=dev)); G.bias_act_backward(dy,pre,G.ACT_GELU,b)
agg=SmallAggregator(260,0,1,dev); agg.slot().normal_(); agg.allreduce(1.0)
X=torch.randn(1000000,256,device=dev,dtype=torch.bfloat16); yy=(torch.rand(1000000,device=dev)<0.5).float(); K8.logistic_grad(X,yy,torch.zeros(257,device=dev))
torch.cuda.synchronize()" > gpurun_out/ncu_late.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_late.log
ls -la gpurun_out/prof_late.ncu-rep
