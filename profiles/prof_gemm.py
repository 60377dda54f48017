"""ncu target: the tcgen05 projection kernel at the teacher shape (2.45M x 256 x 256, C = A W^T) and
at the student shapes (4096 x 2048 x 2048 forward / dX / dW)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "teacher"
if which == "teacher":
    a = torch.randn(2449029, 256, device=dev); w = torch.randn(256, 256, device=dev)
    out = torch.empty(2449029, 256, device=dev)
    for _ in range(3):
        ops.gemm(a, w, trans_b=True, out=out, impl=2)
else:
    x = torch.randn(4096, 2048, device=dev); w = torch.randn(2048, 2048, device=dev)
    dy = torch.randn(4096, 2048, device=dev)
    for _ in range(3):
        ops.gemm(x, w, trans_b=True, impl=2)          # forward  (K-major, K-major)
        ops.gemm(dy, w, impl=2)                        # dX       (K-major, MN-major)
        ops.gemm(dy, x, trans_a=True, impl=2)          # dW       (MN-major, MN-major)
torch.cuda.synchronize()
print("done")
