"""Aggregation kernel at the three layer shapes of the products forward (q24 rows), 5 timed launches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
from glnn_b200.workloads import dataset_graph
dev = torch.device("cuda:0")
g = dataset_graph("ogbn-products", device=dev)
n, e = g.num_nodes(), g.num_edges()
def ms(fn, iters=5):
    for _ in range(2): fn()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    t.record(); torch.cuda.synchronize()
    return s.elapsed_time(t) / iters
for d in (256, 100, 48):
    x = torch.randn(n, d, device=dev)
    q = ops.quantize_q24(x)
    pl = ops.new_planes(n, d, dev)
    t = ms(lambda: ops.spmm(g.indptr, g.indices, q, out_planes=pl, self_add=True, mean_plus_one=True))
    print(json.dumps(dict(d=d, row_bytes=q.ldq, ms=round(t, 3), gather_GBps=round(e * q.ldq / t / 1e6, 1))), flush=True)
