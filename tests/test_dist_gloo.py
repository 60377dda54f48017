"""CPU, world_size 2 and 3 over gloo: the dst-row sharding, column relabelling into the padded
replica layout and the per-layer exchange plan of dist_teacher reproduce the single-process
forward.  The layer math is a torch test double (tests/helpers.TorchKernels); on the GPU box the
same host logic drives the CUDA kernels (tests/test_gpu_dist.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    rng = np.random.default_rng(0)
    n = 500
    src = rng.integers(0, n, 4000)
    dst = np.floor(n * rng.random(4000) ** 2).astype(np.int64)
    torch.manual_seed(0)
    dims = [20, 32, 32, 7]
    layers = [(torch.randn(dims[i + 1], dims[i]) * 0.3, torch.randn(dims[i + 1])) for i in range(3)]
    norms = [(torch.rand(32) + 0.5, torch.randn(32)) for _ in range(2)]
    x = torch.randn(n, 20)
    return n, src, dst, layers, norms, x


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    from helpers import TorchKernels
    from glnn_b200 import dist_teacher as DT
    from glnn_b200.graph import graph
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, src, dst, layers, norms, x = _problem()
    g = graph((src, dst), num_nodes=n)
    sg = DT.ShardedGraph(g, rank, world)
    out = DT.sage_forward_sharded(sg, sg.to_padded(x), layers, norms, kernels=TorchKernels)
    q.put((rank, sg.from_padded(out).numpy(), sg.cuts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_forward_equals_single_process(world):
    import glnn_oracle as O
    n, src, dst, layers, norms, x = _problem()
    indptr, indices = O.csr_from_edges(src, dst, n)
    bn = [(s, sh, torch.zeros(32), torch.ones(32) - 1e-5) for s, sh in norms]
    want = torch.log_softmax(O.sage_inference(indptr, indices, x, layers, bn), 1).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, cuts in results:
        assert np.abs(got - want).max() < 1e-4, rank  # every rank ends with the full result
        assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == world + 1


def test_nnz_balanced_cuts():
    from glnn_b200.dist_teacher import nnz_balanced_cuts
    deg = torch.tensor([1000] + [1] * 999)
    indptr = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)])
    cuts = nnz_balanced_cuts(indptr, 4, row_cost=1)
    loads = [int(indptr[cuts[i + 1]] - indptr[cuts[i]]) + cuts[i + 1] - cuts[i] for i in range(4)]
    assert cuts[0] == 0 and cuts[-1] == 1000 and cuts == sorted(cuts)
    assert max(loads) <= 1001 + 2 * 500  # the hub row dominates one shard, the rest stay even
    even = nnz_balanced_cuts(indptr, 4, row_cost=16)
    sizes = [even[i + 1] - even[i] for i in range(4)]
    assert max(sizes) - min(sizes) < 80   # row-weighted cuts keep the slabs (and the padding) even
