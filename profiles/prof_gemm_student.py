"""ncu target: the tcgen05 planes projection at the student's shapes (bs 4096, H 2048):
short-K forward (K = 104, epilogue-bound) and the square forward (K = 2048)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
dev = torch.device("cuda:0")
x0 = ops.split_planes(torch.randn(4096, 100, device=dev)); w0 = ops.split_planes(torch.randn(2048, 100, device=dev))
x1 = ops.split_planes(torch.randn(4096, 2048, device=dev)); w1 = ops.split_planes(torch.randn(2048, 2048, device=dev))
b = torch.randn(2048, device=dev)
out = torch.empty(4096, 2048, device=dev)
for _ in range(3):
    ops.gemm_planes(x0, w0, trans_b=True, out=out, bias=b)
    ops.gemm_planes(x1, w1, trans_b=True, out=out, bias=b)
torch.cuda.synchronize()
print("done")
