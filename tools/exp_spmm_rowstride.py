"""Narrow-row gathers run below the DRAM peak (round 1: 4.2 TB/s at 160-byte rows, 5.1 TB/s at 320).
This experiment times the products-graph aggregation for the two narrow widths of the forward (48 and
100 columns) with different row formats / strides, to see whether DRAM access granularity (rows that
straddle 64-byte atoms) explains the gap:  q24 at its natural stride, q24 padded to a 64-byte multiple,
fp32 at its natural stride, fp32 padded.  One JSON line per variant."""
import json
import os
import sys

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


for d in (48, 100):
    x = torch.randn(n, d, device=dev)
    out = torch.empty(n, (d + 7) // 8 * 8, device=dev)
    variants = []
    nat = ops.Q24.row_bytes(d)
    for ldq in sorted({nat, (nat + 63) // 64 * 64, (nat + 127) // 128 * 128}):
        buf = torch.zeros(n, ldq, dtype=torch.uint8, device=dev)
        xq = ops.quantize_q24(x, out=ops.Q24(buf, d))
        variants.append((f"q24 stride {ldq}", xq, ldq))
    for ldx in sorted({d, (d * 4 + 63) // 64 * 16, (d * 4 + 127) // 128 * 32}):
        xf = torch.zeros(n, ldx, device=dev)
        xf[:, :d] = x
        variants.append((f"fp32 stride {4 * ldx}", xf[:, :d], 4 * ldx))
    for name, mat, stride in variants:
        o = out if isinstance(mat, ops.Q24) else out[:, :d]
        t = ms(lambda: ops.spmm(g.indptr, g.indices, mat, out=o, self_add=True, mean_plus_one=True))
        print(json.dumps(dict(d=d, variant=name, ms=round(t, 3), row_stride=stride,
                              stride_TBps=round(e * stride / t / 1e9, 3))), flush=True)
    del x, out, variants
