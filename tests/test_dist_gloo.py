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


def _worker(rank, world, port, q, chunks=1, local_feats=False):
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
    sg = DT.ShardedGraph(g, rank, world, chunks=chunks)
    if local_feats:  # every rank ships only its own rows; the input replica is exchanged too
        out = DT.sage_forward_sharded(sg, None, layers, norms, kernels=TorchKernels,
                                      feats_local=x[sg.r0:sg.r0 + sg.rows].contiguous())
    else:
        out = DT.sage_forward_sharded(sg, sg.to_padded(x), layers, norms, kernels=TorchKernels)
    assert sg.total_rows == world * sg.rc * chunks and sg.rows_max == sg.rc * chunks
    q.put((rank, sg.from_padded(out).numpy(), sg.cuts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,chunks,local_feats", [(2, 1, False), (3, 1, False), (2, 3, False),
                                                      (3, 4, False), (2, 3, True), (3, 1, True)])
def test_sharded_forward_equals_single_process(world, chunks, local_feats):
    import glnn_oracle as O
    n, src, dst, layers, norms, x = _problem()
    indptr, indices = O.csr_from_edges(src, dst, n)
    bn = [(s, sh, torch.zeros(32), torch.ones(32) - 1e-5) for s, sh in norms]
    want = torch.log_softmax(O.sage_inference(indptr, indices, x, layers, bn), 1).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world + 7 * chunks + 31 * int(local_feats) + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, chunks, local_feats))
             for r in range(world)]
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


def _student_problem(norm):
    gen = torch.Generator().manual_seed(3)
    f, h, c, L, n, bs, nb = 12, 16, 5, 3, 200, 24, 4
    p = {}
    dims = [f, h, h, c]
    for l in range(L):
        p[f"layers.{l}.weight"] = torch.randn(dims[l + 1], dims[l], generator=gen, dtype=torch.float64) * 0.3
        p[f"layers.{l}.bias"] = torch.randn(dims[l + 1], generator=gen, dtype=torch.float64) * 0.1
    if norm == "batch":
        for l in range(L - 1):
            p[f"norms.{l}.weight"] = torch.rand(h, generator=gen, dtype=torch.float64) + 0.5
            p[f"norms.{l}.bias"] = torch.randn(h, generator=gen, dtype=torch.float64) * 0.1
            p[f"norms.{l}.running_mean"] = torch.zeros(h, dtype=torch.float64)
            p[f"norms.{l}.running_var"] = torch.ones(h, dtype=torch.float64)
            p[f"norms.{l}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    x = torch.randn(n, f, generator=gen, dtype=torch.float64)
    soft = torch.log_softmax(torch.randn(n, c, generator=gen, dtype=torch.float64), 1)
    hard = torch.randint(0, c, (n,), generator=gen)
    idx = torch.randperm(n, generator=gen)[: nb * bs].view(nb, bs)
    return p, x, soft, hard, idx, L


def _student_worker(rank, world, port, q, norm):
    for pth in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, pth)
    import glnn_oracle as O
    from helpers import dp_student_pass
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p, x, soft, hard, idx, L = _student_problem(norm)
    st = O.init_adam_state(p)
    l1 = dp_student_pass(dist, rank, world, p, st, x, hard, "nll", idx, 0.3, L, norm, 0.01, 5e-4)
    l2 = dp_student_pass(dist, rank, world, p, st, x, soft, "kl", idx, 0.7, L, norm, 0.01, 5e-4)
    q.put((rank, l1, l2, {k: v.numpy().copy() for k, v in p.items()},
           {k: v["exp_avg_sq"].numpy().copy() for k, v in st.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,norm", [(2, "batch"), (3, "batch"), (2, "none")])
def test_student_data_parallel_protocol_equals_single_process(world, norm):
    """The decomposition that glnn_mlp_train_pass_dp fuses into its kernels (global-batch BatchNorm
    from per-rank partials, dgamma/dbeta contributed once, 1/B_global loss scaling, slice-owned Adam
    with parameter all-gather), restated in fp64 torch and run over gloo, reproduces the oracle's
    single-process train_mini_batch (train_and_eval.py:59-86) to fp64 rounding."""
    import glnn_oracle as O
    p, x, soft, hard, idx, L = _student_problem(norm)
    st = O.init_adam_state(p)
    w1 = O.train_mini_batch(p, st, x, hard, "nll", idx.shape[1], idx, 0.3, L, norm, 0.0, 0.01, 5e-4)
    w2 = O.train_mini_batch(p, st, x, soft, "kl", idx.shape[1], idx, 0.7, L, norm, 0.0, 0.01, 5e-4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + world + (os.getpid() % 150)
    procs = [ctx.Process(target=_student_worker, args=(r, world, port, q, norm)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, l1, l2, got, got_v in results:
        assert abs(l1 - w1) < 1e-12 and abs(l2 - w2) < 1e-12, (rank, l1, w1, l2, w2)
        for k, v in p.items():
            assert np.allclose(got[k].astype(np.float64), v.double().numpy(), rtol=1e-9, atol=1e-12), (rank, k)
        for k, v in st.items():
            assert np.allclose(got_v[k], v["exp_avg_sq"].numpy(), rtol=1e-9, atol=1e-30), (rank, k)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_two_pass_split_partitions_the_edges_and_the_push_plan_is_consistent(world):
    """Host logic of the two-pass exchange (dist_teacher.py): for every rank the owner-split CSRs hold
    every edge of the rank's slice exactly once, in CSR order, early edges = sources owned by the early
    owners (the explicit self edge included); rank r is an early receiver of s exactly when s is an
    early owner of r; every ordered pair of distinct ranks is served exactly once (early or late)."""
    import torch
    from glnn_b200 import dist_teacher as DT
    from glnn_b200.workloads import synthetic_graph
    g = synthetic_graph(3000, 20000, mirror=True, self_loops=False, device="cpu", seed=2)
    shards = [DT.ShardedGraph(g, r, world, chunks=2) for r in range(world)]
    for r, sg in enumerate(shards):
        early = sg.early_owners()
        assert early[0] == r and len(early) == max(1, world // 2) and len(set(early)) == len(early)
        for s in range(world):
            assert (r in shards[s].early_receivers()) == (s in early and s != r)
        (pe, ie), (pl, il) = sg.split_by_owner()
        assert int(pe[-1]) + int(pl[-1]) == sg.indices.numel()
        owner = lambda idx: (idx.long() // sg.rc) % world
        assert set(owner(ie).tolist()) <= set(early)
        assert not (set(owner(il).tolist()) & set(early))
        p = sg.indptr.long()
        for v in (0, 1, sg.rows // 2, sg.rows - 1):
            row = sg.indices[p[v]:p[v + 1]]
            e_part, l_part = ie[pe[v]:pe[v + 1]], il[pl[v]:pl[v + 1]]
            m = torch.tensor([int(o) in early for o in owner(row)])
            assert torch.equal(row[m], e_part) and torch.equal(row[~m], l_part)   # order kept
            # the self edge (own slot in the replica) is always in the early half
            self_pid = sg.pad_ids[sg.r0 + v]
            assert int((e_part.long() == self_pid).sum()) >= 1
    # every ordered pair (sender, receiver) appears exactly once in the sender's push plan
    for s, sg in enumerate(shards):
        early_r = sg.early_receivers()
        late_r = [(s + i) % world for i in range(1, world) if (s + i) % world not in early_r]
        assert sorted(early_r + late_r) == [x for x in range(world) if x != s]
