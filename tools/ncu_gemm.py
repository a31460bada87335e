"""A few representative GEMM launches between cudaProfilerStart/Stop for `ncu --set full --profile-from-start off`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from counting_detr_b200 import _lib as L
dev = "cuda"
torch.manual_seed(0)
shapes = [  # (M, N, K, kind)  layer4 conv2 as GEMM; encoder FFN linear1 (plain, and with residual + ReLU mask); layer1 conv3
    (16384, 512, 4608, "conv3x3"), (16384, 1024, 256, "ffn1"), (16384, 1024, 256, "ffn1+res+mask"), (262144, 256, 64, "conv1x1+res")]
ops = []
for M, N, K, kind in shapes:
    A = L.to_split(torch.randn(M, K, device=dev)); B = L.to_split(torch.randn(N, K, device=dev))
    out = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16)
    res = L.to_split(torch.randn(M, N, device=dev)) if "res" in kind else None
    mask = L.to_split(torch.randn(M, N, device=dev)) if "mask" in kind else None
    bias = torch.randn(N, device=dev)
    ops.append((A, B, M, N, K, out, res, mask, bias))
def run():
    for A, B, M, N, K, out, res, mask, bias in ops:
        L.gemm(A, B, M, N, K, out_split=out, add_split=res, mask=mask, bias=bias, relu=mask is None)
for _ in range(3): run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
