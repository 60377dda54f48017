"""Graph construction on the device: glnn_csr_from_coo (degree counts + scan + 8-bit LSD radix passes,
csrc/csr_build.cu) against the same result assembled from library ops (torch.sort(stable) + gather +
bincount + cumsum, what round 1 used), on the products-sized edge list (123.7M edges, 2.45M nodes)
and the arxiv-sized one; plus glnn_csr_subgraph against the edge-list route for a 80 % node subset
(the inductive split).  One JSON line per variant; algorithmic bytes = read src + dst (int64), write
indices + indptr + out_deg."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import ops
from glnn_b200.graph import CSRGraph
from glnn_b200.workloads import SHAPES, synthetic_edges

dev = torch.device("cuda:0")
# --tile-sweep: only the products-sized glnn_csr_from_coo, for the radix tile chosen by GLNN_CSR_ITEMS
#               (one process per value: the library reads it once), with and without out-degrees
# --once:       one products-sized call and nothing else (for an ncu launch list / --set full capture)
MODE = sys.argv[1] if len(sys.argv) > 1 else ""


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


def library_route(src, dst, n):
    order = torch.sort(dst, stable=True).indices
    indices = src[order].to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=src.device)
    torch.cumsum(torch.bincount(dst, minlength=n), 0, out=indptr[1:])
    return indptr.to(torch.int32), indices, torch.bincount(src, minlength=n)


if MODE in ("--tile-sweep", "--once"):
    s = SHAPES["ogbn-products"]
    n = s["n"]
    src, dst = synthetic_edges(n, s["e_raw"], True, s["self_loops"], dev, 0)
    if MODE == "--once":
        ops.csr_from_coo(src, dst, n)
        torch.cuda.synchronize()
        sys.exit(0)
    ref = library_route(src, dst, n)
    for want in (True, False):
        got = ops.csr_from_coo(src, dst, n, want_out_deg=want)
        same = torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and (not want or torch.equal(got[2], ref[2]))
        t = ms(lambda: ops.csr_from_coo(src, dst, n, want_out_deg=want))
        print(json.dumps({"exp": "csr_from_coo tile sweep", "items_per_thread": os.environ.get("GLNN_CSR_ITEMS", "8"),
                          "out_degrees": want, "identical": same, "ms": round(t, 3)}), flush=True)
    sys.exit(0)

for name in ("ogbn-arxiv", "ogbn-products"):
    s = SHAPES[name]
    n = s["n"]
    src, dst = synthetic_edges(n, s["e_raw"], True, s["self_loops"], dev, 0)
    e = src.numel()
    a = ops.csr_from_coo(src, dst, n)
    b = library_route(src, dst, n)
    same = all(torch.equal(x, y) for x, y in zip(a, b))
    del a, b
    alg = 16 * e + 4 * e + 4 * (n + 1) + 8 * n
    t_own = ms(lambda: ops.csr_from_coo(src, dst, n))
    t_lib = ms(lambda: library_route(src, dst, n))
    print(json.dumps({"exp": "csr_from_coo", "graph": name, "nodes": n, "edges": e, "identical": same,
                      "own_kernels_ms": round(t_own, 3), "library_ops_ms": round(t_lib, 3),
                      "algorithmic_GB": round(alg / 1e9, 3),
                      "own_GBps_algorithmic": round(alg / t_own / 1e6, 1)}), flush=True)
    g = CSRGraph.from_edges(src, dst, n)
    del src, dst
    nodes = torch.randperm(n, device=dev)[: int(0.8 * n)]

    def edge_list_route():
        relabel = torch.full((n,), -1, dtype=torch.int64, device=dev)
        relabel[nodes] = torch.arange(nodes.numel(), device=dev)
        s0, d0 = g.edges()
        s1, d1 = relabel[s0], relabel[d0]
        keep = (s1 >= 0) & (d1 >= 0)
        return library_route(s1[keep], d1[keep], nodes.numel())

    sub = g.subgraph(nodes)
    ref = edge_list_route()
    same = torch.equal(sub.indptr, ref[0]) and torch.equal(sub.indices, ref[1]) and \
        torch.equal(sub.out_degrees(), ref[2])
    del ref
    t_own = ms(lambda: g.subgraph(nodes), iters=3)
    t_lib = ms(edge_list_route, iters=3)
    print(json.dumps({"exp": "csr_subgraph", "graph": name, "kept_nodes": nodes.numel(),
                      "kept_edges": sub.num_edges(), "identical": same, "own_kernels_ms": round(t_own, 3),
                      "library_ops_ms": round(t_lib, 3)}), flush=True)
    del g, sub
    torch.cuda.empty_cache()
