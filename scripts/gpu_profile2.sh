#!/bin/bash
# 1-GPU: ncu --set full captures of the hand-written kernels of the ResNet-50 local step (BN trees, pooling, stem
# transforms, gradient sink) and of the GEMMs with the TMA-store epilogue.  .ncu-rep files land in gpurun_out/.
mkdir -p gpurun_out
export PYTHONPATH=.
echo "== ncu full: resnet step kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bn_|maxpool|multi_accum|image_normalize|stem_" -c 14 -o gpurun_out/prof_resnet -f \
  python -c "
import torch
from vantage6_b200.ops.bn import FusedBatchNormAct
from vantage6_b200.ops.pool import MaxPool3x3s2, stem_s2d
from vantage6_b200.ops.optim import multi_accumulate
from vantage6_b200.models.resnet import _MEAN, _STD
dev=torch.device('cuda',0); cl=torch.channels_last
def run(N,C,H,W,res):
    x=torch.randn(N,C,H,W,device=dev).to(torch.bfloat16).contiguous(memory_format=cl).requires_grad_()
    r=torch.randn(N,C,H,W,device=dev).to(torch.bfloat16).contiguous(memory_format=cl).requires_grad_() if res else None
    bn=FusedBatchNormAct(C).to(dev); y=bn(x,r); y.backward(torch.ones_like(y))
run(64,256,56,56,True)      # 103 MB activation: HBM-bound
run(64,1024,14,14,False)    # 25.7 MB: L2-resident
x=torch.randn(64,64,112,112,device=dev).to(torch.bfloat16).contiguous(memory_format=cl).requires_grad_()
y=MaxPool3x3s2()(x); y.backward(torch.ones_like(y))
img=torch.randint(0,256,(64,3,224,224),dtype=torch.uint8,device=dev)
conv=torch.nn.Conv2d(3,64,7,stride=2,padding=3,bias=False).to(dev).to(memory_format=cl)
o=stem_s2d(img,conv,_MEAN,_STD); o.backward(torch.ones_like(o))
dst=torch.zeros(25_000_000,device=dev); gs=[(torch.randn(2359296,device=dev).to(torch.bfloat16), i*2359296) for i in range(8)]
multi_accumulate(dst,gs)
torch.cuda.synchronize()" > gpurun_out/ncu_resnet.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_resnet.log
echo "== ncu full: gemm (TMA-store epilogue), 1-CTA small-K and 2-CTA large"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|gemm2_bf16_kernel" -s 2 -c 2 -o gpurun_out/prof_gemm2 -f \
  python -c "
import torch; from vantage6_b200.ops import gemm as G
a=torch.randn(4096,768,device='cuda',dtype=torch.bfloat16); w=torch.randn(3072,768,device='cuda',dtype=torch.bfloat16)
A=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16); W=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16)
G.gemm_bf16(a,w,variant='1cta'); G.gemm_bf16(A,W,variant='2cta'); G.gemm_bf16(a,w,variant='1cta'); G.gemm_bf16(A,W,variant='2cta')
torch.cuda.synchronize()" > gpurun_out/ncu_gemm2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_gemm2.log
ls -la gpurun_out/*.ncu-rep
