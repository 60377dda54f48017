"""ncu target: the student's short-K forward projection (4096 x 2048 x 100, bias, fp32 out): 256
CTAs of one 128 x 256 tile, same per-CTA timeline as the small students' projections."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
dev = torch.device("cuda:0")
x = ops.split_planes(torch.randn(4096, 100, device=dev)); w = ops.split_planes(torch.randn(2048, 100, device=dev))
b = torch.randn(2048, device=dev)
out = torch.empty(4096, 2048, device=dev)
for _ in range(4):
    ops.gemm_planes(x, w, trans_b=True, out=out, bias=b)
torch.cuda.synchronize()
print("done")
