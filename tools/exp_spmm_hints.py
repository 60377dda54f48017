"""Experiment: effect of L2 eviction hints (hot_below) and q24 row formats on the aggregation kernel,
ogbn-products-shaped graph.  Prints one line per (format, width, L2 budget)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
from glnn_b200.workloads import dataset_graph

dev = torch.device("cuda:0")
g = dataset_graph("ogbn-products", device=dev)
n, e = g.num_nodes(), g.num_edges()


def ms(fn, iters=5):
    for _ in range(2):
        fn()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / iters


res = []
for d in (256, 100, 48):
    x = torch.randn(n, d, device=dev)
    q = ops.quantize_q24(x)
    pl = ops.new_planes(n, d, dev)
    yf = torch.empty(n, (d + 7) // 8 * 8, device=dev)
    for fmt, src, rb in (("f32", x, 4 * d), ("q24", q, q.ldq)):
        for mb in (0, 20, 40, 60, 80, 100, 120):
            hot = int(min(n, mb * 1e6 / rb))
            t = ms(lambda: ops.spmm(g.indptr, g.indices, src, out_planes=pl, self_add=True,
                                    mean_plus_one=True, hot_below=hot))
            line = dict(d=d, fmt=fmt, row_bytes=rb, hot_mb=mb, hot_rows=hot, ms=round(t, 3),
                        gather_GBps=round(e * rb / t / 1e6, 1))
            print(json.dumps(line), flush=True)
            res.append(line)
    del x, q, pl, yf
json.dump(res, open("gpurun_out/exp_spmm_hints.json", "w"), indent=1)
