"""On-device graph construction (COO -> CSR sort/scan and subgraph(idx_obs)):
glnn_csr_from_coo / glnn_csr_subgraph against the oracle's host builder (numpy stable argsort, the
order DGL keeps for dgl.graph((src, dst)), dataloader.py:78) -- index work, so bit-exact."""
import numpy as np
import pytest
import torch

import glnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _edges(n, e, seed, skew=True):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, e)
    dst = np.minimum((n * rng.random(e) ** 2).astype(np.int64), n - 1) if skew else rng.integers(0, n, e)
    return src.astype(np.int64), dst.astype(np.int64)


# nodes, edges: one / two / three / four radix passes, ragged last tile, exact tile, empty edge list,
# a single node, isolated nodes (n >> e), heavy multi-edges (e >> n)
CASES = [(1, 10), (2, 50), (5, 0), (256, 3000), (257, 5000), (300, 2048), (300, 2049), (300, 2047),
         (70000, 300000), (65537, 1000003), (40, 200000), (3000000, 1000), (2 ** 24 + 5, 100000)]


@pytest.mark.parametrize("n,e", CASES)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_csr_from_coo_equals_host_stable_sort(dev, n, e, dtype):
    from glnn_b200 import ops
    src, dst = _edges(n, e, seed=n + e)
    want_ptr, want_idx = O.csr_from_edges(src, dst, n)
    indptr, indices, out_deg = ops.csr_from_coo(torch.from_numpy(src).to(dev, dtype),
                                                torch.from_numpy(dst).to(dev, dtype), n)
    assert indptr.dtype == torch.int32 and indices.dtype == torch.int32 and out_deg.dtype == torch.int64
    assert np.array_equal(indptr.cpu().numpy().astype(np.int64), want_ptr)
    assert np.array_equal(indices.cpu().numpy().astype(np.int64), want_idx)   # order inside rows too
    assert np.array_equal(out_deg.cpu().numpy(), np.bincount(src, minlength=n))


def test_graph_constructor_uses_the_device_builder_and_matches_the_host_one(dev):
    from glnn_b200.graph import CSRGraph, graph
    src, dst = _edges(5000, 80000, seed=3)
    g_host = graph((src, dst), num_nodes=5000)
    g_dev = CSRGraph.from_edges(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), 5000)
    assert g_dev.indices.is_cuda
    assert torch.equal(g_dev.indptr.cpu(), g_host.indptr) and torch.equal(g_dev.indices.cpu(), g_host.indices)
    assert torch.equal(g_dev.out_degrees().cpu(), g_host.out_degrees())
    assert torch.equal(g_dev.in_degrees().cpu(), g_host.in_degrees())
    # num_nodes inferred from the ids
    g2 = CSRGraph.from_edges(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev))
    assert g2.num_nodes() == int(max(src.max(), dst.max())) + 1


def test_out_of_range_ids_raise(dev):
    from glnn_b200 import ops
    src = torch.tensor([0, 1, 7], device=dev)
    dst = torch.tensor([1, 2, 0], device=dev)
    with pytest.raises(ValueError):
        ops.csr_from_coo(src, dst, 5)
    with pytest.raises(ValueError):
        ops.csr_from_coo(torch.tensor([0, -1], device=dev), torch.tensor([1, 1], device=dev), 5)


@pytest.mark.parametrize("n,e,frac", [(4000, 60000, 0.5), (4000, 60000, 1.0), (300, 5000, 0.05),
                                      (50000, 400000, 0.3), (10, 0, 0.5)])
def test_subgraph_equals_host_subgraph(dev, n, e, frac):
    """g.subgraph(idx_obs) (train_and_eval.py:324): nodes relabelled in the GIVEN (shuffled) order,
    rows keep the order of their kept edges; compared with the host implementation edge for edge."""
    from glnn_b200.graph import graph
    src, dst = _edges(n, e, seed=n + 1)
    g_host = graph((src, dst), num_nodes=n)
    g_host.ndata["feat"] = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
    g_dev = g_host.to(dev)
    gen = torch.Generator().manual_seed(5)
    nodes = torch.randperm(n, generator=gen)[: max(1, int(n * frac))]
    want = g_host.subgraph(nodes)
    got = g_dev.subgraph(nodes.to(dev))
    assert got.num_nodes() == want.num_nodes() and got.indices.is_cuda
    assert torch.equal(got.indptr.cpu().to(torch.int64), want.indptr.to(torch.int64))
    assert torch.equal(got.indices.cpu(), want.indices)
    assert torch.equal(got.out_degrees().cpu(), want.out_degrees())
    assert torch.equal(got.ndata["feat"].cpu(), want.ndata["feat"])


def test_subgraph_of_nothing_and_of_isolated_nodes(dev):
    from glnn_b200.graph import graph
    g = graph((np.array([0, 1, 2]), np.array([1, 2, 0])), num_nodes=6).to(dev)
    sub = g.subgraph(torch.tensor([4, 5, 3], device=dev))           # isolated nodes only
    assert sub.num_edges() == 0 and sub.indptr.cpu().tolist() == [0, 0, 0, 0]
    sub = g.subgraph(torch.tensor([2, 0], device=dev))              # keeps the edge 2 -> 0 only
    assert sub.indptr.cpu().tolist() == [0, 0, 1] and sub.indices.cpu().tolist() == [0]


def test_full_size_products_graph_properties(dev):
    """BASELINE's largest graph (2.45M nodes, 123.7M edges, hubs of ~80k in-edges) through the device
    builder, checked by size-independent properties: degree sums, per-row checksums and -- on the
    biggest hubs and a random sample of rows -- the exact order of the row against the edge list."""
    from glnn_b200.graph import CSRGraph
    from glnn_b200.workloads import SHAPES, synthetic_edges
    s = SHAPES["ogbn-products"]
    n = s["n"]
    src, dst = synthetic_edges(n, s["e_raw"], True, s["self_loops"], dev, 0)
    e = src.numel()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    g = CSRGraph.from_edges(src, dst, n)
    t1.record()
    torch.cuda.synchronize()
    print(f"\ncsr_from_coo: {e} edges, {n} nodes in {t0.elapsed_time(t1):.1f} ms")
    ptr64 = g.indptr.to(torch.int64)
    assert int(ptr64[0]) == 0 and int(ptr64[-1]) == e
    deg = ptr64[1:] - ptr64[:-1]
    assert torch.equal(deg, torch.bincount(dst, minlength=n))
    assert torch.equal(g.out_degrees(), torch.bincount(src, minlength=n))
    assert int(g.indices.min()) >= 0 and int(g.indices.max()) < n
    # per-row checksums: sum and sum of squares of the source ids of every row
    rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    idx = g.indices.to(torch.int64)
    for f in (lambda x: x, lambda x: (x * x) % 1000003):
        want = torch.zeros(n, dtype=torch.int64, device=dev).index_add_(0, dst, f(src))
        got = torch.zeros(n, dtype=torch.int64, device=dev).index_add_(0, rows, f(idx))
        assert torch.equal(got, want)
    del rows, idx
    # exact row contents, in input order
    gen = torch.Generator().manual_seed(0)
    sample = torch.cat([torch.arange(4), torch.randint(0, n, (60,), generator=gen)]).tolist()
    for v in sample:
        want = src[dst == v].to(torch.int32)
        assert torch.equal(g.indices[int(ptr64[v]):int(ptr64[v + 1])], want), v
