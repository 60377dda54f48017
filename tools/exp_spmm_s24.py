"""EXPERIMENTAL: layer-1 aggregation of the products forward from sparse (s24) rows vs q24 rows.
Prints the per-row capacity, both timings and the compaction pass.  """
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
d = 256
for zero_frac in (0.52, 0.7, 0.3):
    x = torch.randn(n, d, device=dev)
    x = torch.relu(x - float(torch.quantile(x.flatten()[:1000000], zero_frac)))
    xq = ops.quantize_q24(x); del x
    pl = ops.new_planes(n, d, dev)
    t_c = ms(lambda: ops.compact_s24(xq))
    s = ops.compact_s24(xq)
    t_q = ms(lambda: ops.spmm(g.indptr, g.indices, xq, out_planes=pl, self_add=True, mean_plus_one=True))
    t_s = ms(lambda: ops.spmm(g.indptr, g.indices, xq, out_planes=pl, self_add=True, mean_plus_one=True, s24=s))
    print(json.dumps(dict(zero_frac=zero_frac, cap=int(s.cap.item()), q24_ms=round(t_q, 3), s24_ms=round(t_s, 3),
                          compact_ms=round(t_c, 3))), flush=True)
    del s, xq, pl
