"""ncu target: the student projection at the SMALL shapes of the arxiv students (bs 512, H 256):
forward 512 x 256 x 128 and the weight gradient 256 x 256 x 512 (both operands MN-major), where the
kernel shows an 11-22 us floor (profiles/r2_launches_student_h256_bs512.csv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
dev = torch.device("cuda:0")
x = ops.split_planes(torch.randn(512, 128, device=dev)); w = ops.split_planes(torch.randn(256, 128, device=dev))
dz = ops.split_planes(torch.randn(512, 256, device=dev)); a = ops.split_planes(torch.randn(512, 256, device=dev))
b = torch.randn(256, device=dev)
out = torch.empty(512, 256, device=dev); dw = torch.empty(256, 256, device=dev)
for _ in range(3):
    ops.gemm_planes(x, w, trans_b=True, out=out, bias=b)
    ops.gemm_planes(dz, a, trans_a=True, out=dw)
torch.cuda.synchronize()
print("done")
