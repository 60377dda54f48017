"""ncu target: the dominant kernel of the teacher forward -- the neighbour aggregation of layer 1 on
the ogbn-products-shaped graph: 256-wide q24 rows (768 B) gathered into bf16 hi/lo planes --
launched 3 times through the C ABI.  `python prof_spmm.py 100` / `48` give the layer-0 / layer-2
shapes (320 B / 160 B rows); `python prof_spmm.py 256 f32` the fp32-row variant of round 1a."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
from glnn_b200.workloads import dataset_graph
d = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fmt = sys.argv[2] if len(sys.argv) > 2 else "q24"
dev = torch.device("cuda:0")
g = dataset_graph("ogbn-products", device=dev)
x = torch.randn(g.num_nodes(), d, device=dev)
src = ops.quantize_q24(x) if fmt == "q24" else x
out = ops.new_planes(g.num_nodes(), d, dev)
for _ in range(3):
    ops.spmm(g.indptr, g.indices, src, out_planes=out, self_add=True, mean_plus_one=True)
torch.cuda.synchronize()
print("done", out.hi[0, 0].item())
