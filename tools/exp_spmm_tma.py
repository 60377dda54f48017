"""A/B for the north_star's "neighbour features pulled through TMA into shared memory" clause: the
256-wide q24 aggregation (768-byte rows, the dominant kernel of the products forward) with
  (a) spmm_csr_kernel<32,1,8>: ld.global.nc.v4 straight into registers, 8 rows in flight per warp, and
  (b) spmm_tma_q24_kernel<STAGES>: one cp.async.bulk (TMA) per neighbour row into a per-warp
      shared-memory ring + mbarrier, STAGES = 4 / 8 rows in flight per warp,
on a products-SIZED graph with UNIFORM random endpoints (2,449,029 nodes, 123.7 M edges, no hubs: the
experimental kernel has no hub path), checked against each other, timed with CUDA events.
Prints one JSON line per variant."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
from glnn_b200.graph import CSRGraph

dev = torch.device("cuda:0")
n, e = 2449029, 123718280
gen = torch.Generator(device=dev).manual_seed(0)
src = torch.randint(0, n, (e,), device=dev, generator=gen)
dst = torch.randint(0, n, (e,), device=dev, generator=gen)
g = CSRGraph.from_edges(src, dst, n)
del src, dst
x = torch.relu(torch.randn(n, 256, device=dev, generator=gen))
xq = ops.quantize_q24(x)
del x


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


ref = ops.new_planes(n, 256, dev)
ops.spmm(g.indptr, g.indices, xq, out_planes=ref, self_add=True, mean_plus_one=True)
t_ld = ms(lambda: ops.spmm(g.indptr, g.indices, xq, out_planes=ref, self_add=True, mean_plus_one=True))
gb = e * 768 / 1e9
print(json.dumps(dict(variant="ld.global.nc -> registers (product kernel)", ms=round(t_ld, 3),
                      gather_TBps=round(gb / t_ld, 3), max_in_degree=int(g.in_degrees().max()))), flush=True)
for stages in (4, 8):
    out = ops.new_planes(n, 256, dev)
    ops.exp_spmm_tma(g.indptr, g.indices, xq, out, stages)
    torch.cuda.synchronize()
    same = bool(torch.equal(out.hi, ref.hi) and torch.equal(out.lo, ref.lo))
    err = float((out.float() - ref.float()).abs().max())
    t = ms(lambda: ops.exp_spmm_tma(g.indptr, g.indices, xq, out, stages))
    print(json.dumps(dict(variant=f"cp.async.bulk -> smem ring, {stages} stages/warp", ms=round(t, 3),
                          gather_TBps=round(gb / t, 3), bit_identical=same, max_abs_diff=err)), flush=True)
    del out
