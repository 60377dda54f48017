"""ncu target: the dominant kernel of the teacher forward (aggregation at d=256 on the
ogbn-products-shaped graph), launched 3 times through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
from glnn_b200.workloads import dataset_graph
d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
g = dataset_graph("ogbn-products", device=dev)
x = torch.randn(g.num_nodes(), d, device=dev)
y = torch.empty_like(x)
for _ in range(3):
    ops.spmm_csr(g.indptr, g.indices, x, out=y, self_add=True, mean_plus_one=True)
torch.cuda.synchronize()
print("done", float(y[0, 0]))
