"""ncu target: the persistent TMA-fed tcgen05 projection (gemm_tall.cu) at the teacher's layer-1
shape (2,449,029 x 256 x 256, planes in, planes out, +bias +BN affine +ReLU), 3 launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
dev = torch.device("cuda:0")
n, k, d = 2449029, int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 256
a = ops.split_planes(torch.randn(n, k, device=dev))
w = ops.split_planes(torch.randn(d, k, device=dev))
b, sc, sh = (torch.randn(d, device=dev) for _ in range(3))
out = ops.new_planes(n, d, dev)
for _ in range(3):
    ops.gemm_planes(a, w, trans_b=True, out_planes=out, bias=b, col_scale=sc, col_shift=sh, relu=1)
torch.cuda.synchronize()
print("done")
