"""GPU tests of teacher TRAINING as a kernel sequence: the train-mode block
[BatchNorm] -> [ReLU] -> [Dropout] and its backward, the loss kernel, the scatter form of the
transposed aggregation, the neighbour-sampling kernels, and whole `train` / `train_sage` steps with
BatchNorm and dropout against fixtures made by the reference's own functions (autograd over the DGL
shim, oracle/make_golden.py).  No autograd runs on the product side."""
import numpy as np
import pytest
import torch

from helpers import load, relerr, relerr_q

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("bn,relu_post,relu_input,p", [(True, True, False, 0.3), (True, False, True, 0.0),
                                                       (False, True, False, 0.5), (False, False, True, 0.4),
                                                       (True, True, False, 0.0), (False, False, False, 0.0)])
@pytest.mark.parametrize("n,d", [(1000, 70), (77, 256), (5000, 33)])
def test_act_block_forward_backward_vs_torch_autograd(dev, n, d, bn, relu_post, relu_input, p):
    """glnn_act_train_fwd_f32 / _bwd_f32 against torch autograd in fp64 on the CPU, with an injected
    keep-mask: output, running statistics, dX, dgamma, dbeta and the column sums (bias gradient)."""
    from glnn_b200 import ops
    gen = torch.Generator().manual_seed(n + d)
    x = torch.randn(n, d, generator=gen) * 2 + 0.5
    if relu_input:
        x = torch.relu(x)
    mask = (torch.rand(n, d, generator=gen) >= p).to(torch.uint8) if p > 0 else None
    dy = torch.randn(n, d, generator=gen)
    norm = torch.nn.BatchNorm1d(d) if bn else None
    if bn:
        with torch.no_grad():
            norm.weight.copy_(torch.rand(d, generator=gen) + 0.5)
            norm.bias.copy_(torch.randn(d, generator=gen) * 0.3)
            norm.running_mean.copy_(torch.randn(d, generator=gen))
            norm.running_var.copy_(torch.rand(d, generator=gen) + 0.5)
    # reference: fp64 autograd
    import copy
    ref = copy.deepcopy(norm).double().train() if bn else None
    xr = x.double().requires_grad_(True)
    t = ref(xr) if bn else xr
    if relu_post:
        t = torch.relu(t)
    if p > 0:
        t = t * mask.double() / (1 - p)
    t.backward(dy.double())
    want_dx = xr.grad.clone()
    if relu_input:
        want_dx = want_dx * (x > 0).double()
    # device
    norm_d = copy.deepcopy(norm).to(dev).train() if bn else None
    blk = ops.ActBlock(x.to(dev), bn=norm_d, relu_post=relu_post, relu_input=relu_input, p_drop=p,
                       keep_mask=None if mask is None else mask.to(dev))
    y = blk.forward()
    assert relerr(y.cpu(), t.detach()) < 2e-6
    dx, dg, db, dbias = blk.backward(dy.to(dev))
    assert relerr(dx.cpu(), want_dx) < 1e-5
    # column sums of dX (the bias gradient): behind BatchNorm they are mathematically zero, so the
    # bound is relative to the column's absolute mass, not to the (vanishing) exact value
    assert float((dbias.cpu().double() - want_dx.sum(0)).abs().max()) < 1e-5 * float(want_dx.abs().sum(0).max())
    if bn:
        assert relerr(norm_d.running_mean.cpu(), ref.running_mean) < 1e-6
        assert relerr(norm_d.running_var.cpu(), ref.running_var) < 1e-5
        assert int(norm_d.num_batches_tracked) == 1
        assert relerr(dg.cpu(), ref.weight.grad) < 1e-5
        assert relerr(db.cpu(), ref.bias.grad) < 1e-5


def test_act_block_device_dropout_stream_is_consistent_and_unbiased(dev):
    """Without a mask the keep decision comes from a counter-based hash: the backward must see the
    same mask as the forward, and the keep rate must be 1 - p."""
    from glnn_b200 import ops
    x = torch.ones(4000, 64, device=dev)
    blk = ops.ActBlock(x, p_drop=0.3, seed=1234)
    y = blk.forward()
    kept = y != 0
    assert abs(float(kept.float().mean()) - 0.7) < 0.01
    assert torch.allclose(y[kept], torch.full_like(y[kept], 1 / 0.7))
    dx, _, _, _ = blk.backward(torch.ones_like(x))
    assert torch.equal(dx != 0, kept)
    y2 = ops.ActBlock(x, p_drop=0.3, seed=1235).forward()
    assert not torch.equal(y2 != 0, kept)
    # keep rate per column and per row band (no structure in the hash)
    assert float((kept.float().mean(0) - 0.7).abs().max()) < 0.04


@pytest.mark.parametrize("subset", [False, True])
def test_nll_loss_grad_vs_torch(dev, subset):
    from glnn_b200 import ops
    gen = torch.Generator().manual_seed(3)
    n, c = 700, 47
    logits = torch.randn(n, c, generator=gen) * 3
    labels_all = torch.randint(0, c, (2000,), generator=gen)
    if subset:   # `train`: loss over idx_train rows of a full logits matrix, labels indexed the same
        rows = torch.randperm(n, generator=gen)[:200]
        lr = logits.double().requires_grad_(True)
        loss = torch.nn.functional.nll_loss(lr.log_softmax(1)[rows], labels_all[:n][rows])
        (0.7 * loss).backward()
        d, got = ops.nll_loss_grad(logits.to(dev), labels_all[:n].to(dev), rows=rows.to(dev), lamb=0.7)
    else:        # `train_sage`: every row, labels picked by output_nodes
        label_rows = torch.randperm(2000, generator=gen)[:n]
        lr = logits.double().requires_grad_(True)
        loss = torch.nn.functional.nll_loss(lr.log_softmax(1), labels_all[label_rows])
        (0.7 * loss).backward()
        d, got = ops.nll_loss_grad(logits.to(dev), labels_all.to(dev), label_rows=label_rows.to(dev), lamb=0.7)
    assert abs(got.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert relerr(d.cpu(), lr.grad) < 1e-5


def test_spmm_scatter_is_the_transpose_of_the_gather(dev):
    from glnn_b200 import ops
    rng = np.random.default_rng(0)
    n_dst, n_src, e, d = 300, 500, 4000, 37
    dst = np.sort(rng.integers(0, n_dst, e))
    src = rng.integers(0, n_src, e)
    indptr = np.zeros(n_dst + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=n_dst), out=indptr[1:])
    A = torch.zeros(n_dst, n_src, dtype=torch.float64)
    for s, t in zip(src, dst):
        A[t, s] += 1.0
    A[:, :n_dst] += torch.eye(n_dst, dtype=torch.float64)        # self term: dst nodes are a prefix
    gen = torch.Generator().manual_seed(1)
    dy = torch.randn(n_dst, d, generator=gen)
    scale = torch.rand(n_dst, generator=gen) + 0.1
    want = A.t() @ (dy.double() * scale.double().unsqueeze(1))
    dx = torch.zeros(n_src, d, device=dev)
    ops.spmm_scatter(torch.from_numpy(indptr).to(dev), torch.from_numpy(src.astype(np.int32)).to(dev),
                     dy.to(dev), scale.to(dev), dx, self_add=True)
    assert relerr(dx.cpu(), want) < 1e-5


def _true_neighbours(g, v):
    p = g.indptr.long()
    return g.indices[p[v]:p[v + 1]].long()


@pytest.mark.parametrize("fanout", [1, 5, 15, 64])
def test_neighbour_sampler_properties(dev, fanout):
    """glnn_sample_neighbors (dgl sample_neighbors(replace=False), train_and_eval.py:179-183): at most
    `fanout` edges per seed, all of them when the in-degree is <= fanout, every sampled edge is a true
    in-edge, NO replacement (a source appears at most as often as it is a true multi-edge), block
    invariants (dst nodes are the prefix of the src nodes, local ids map back to the sampled ids)."""
    from glnn_b200 import teacher_train as TT
    from glnn_b200.workloads import synthetic_graph
    n = 3000
    g = synthetic_graph(n, 40000, mirror=True, self_loops=False, device=dev, seed=5)   # multigraph, hubs
    seeds = torch.randperm(n, generator=torch.Generator().manual_seed(fanout))[:500].to(dev)
    src_nodes, blk = TT.sample_block(g, seeds, fanout, rng_seed=17)
    deg = g.in_degrees()[seeds]
    cnt = (blk.indptr[1:] - blk.indptr[:-1]).long()
    assert torch.equal(cnt, torch.clamp(deg, max=fanout))
    assert blk.n_dst == seeds.numel() and torch.equal(src_nodes[: blk.n_dst], seeds)
    assert src_nodes.unique().numel() == src_nodes.numel()             # a node has ONE local id
    glob = src_nodes[blk.indices.long()]                               # back to global ids
    ip = blk.indptr.long().cpu()
    seeds_c, glob_c = seeds.cpu(), glob.cpu()
    for i in range(0, seeds.numel(), 7):
        v = int(seeds_c[i])
        true = _true_neighbours(g, v).cpu()
        got = glob_c[ip[i]:ip[i + 1]]
        t_ids, t_cnt = true.unique(return_counts=True)
        g_ids, g_cnt = got.unique(return_counts=True)
        pos = torch.searchsorted(t_ids, g_ids)
        assert bool((pos < t_ids.numel()).all()) and torch.equal(t_ids[pos], g_ids)   # subset
        assert bool((g_cnt <= t_cnt[pos]).all())                                      # no replacement
        if true.numel() <= fanout:
            assert torch.equal(got.sort().values, true.sort().values)                 # whole row
    # full neighbourhood (the reference's evaluation sampler / fan-out -1)
    _, full = TT.sample_block(g, seeds, -1)
    assert torch.equal((full.indptr[1:] - full.indptr[:-1]).long(), deg)


def test_neighbour_sampler_is_uniform(dev):
    """Every in-edge of a seed is kept with probability fanout / degree: chi-square over many draws
    of one hub row, and different seeds give different samples."""
    from glnn_b200 import ops
    deg, fanout, draws = 40, 10, 4000
    indptr = torch.tensor([0, deg], dtype=torch.int32, device=dev)
    indices = torch.arange(100, 100 + deg, dtype=torch.int32, device=dev)
    seeds = torch.zeros(1, dtype=torch.int64, device=dev)
    counts = torch.zeros(deg)
    first = None
    for s in range(draws):
        _, src = ops.sample_neighbors(indptr, indices, seeds, fanout, rng_seed=1000 + s)
        src = src.cpu().long() - 100
        assert src.numel() == fanout and src.unique().numel() == fanout
        assert bool((src[1:] > src[:-1]).all())             # CSR order
        counts[src] += 1
        first = src if first is None else first
    expect = draws * fanout / deg
    chi2 = float(((counts - expect) ** 2 / expect).sum())
    assert chi2 < 90, chi2                                  # 39 dof: P(chi2 > 90) ~ 1e-5
    _, again = ops.sample_neighbors(indptr, indices, seeds, fanout, rng_seed=1000)
    assert torch.equal(again.cpu().long() - 100, first)     # same seed -> same sample


def _masks_of(d):
    n_masks = sum(1 for k in d if k.startswith("maskshape."))
    out = []
    for i in range(n_masks):
        shape = tuple(int(x) for x in d[f"maskshape.{i}"])
        bits = np.unpackbits(d[f"mask.{i}"])[: shape[0] * shape[1]].reshape(shape)
        out.append(torch.from_numpy(bits.astype(np.uint8)))
    return out


def _check_final(model, d, q_tol=2e-3, sage=False):
    """Parameters / BN buffers after the last step against the reference's.  With BatchNorm a SAGE
    conv bias sits directly in front of the norm: its mathematical gradient is zero, Adam amplifies
    the rounding noise -- those tensors are skipped."""
    sd = model.state_dict()
    L, bn = int(d["num_layers"]), str(d["norm"]) == "batch"
    for k, v in d.items():
        if not k.startswith("final."):
            continue
        name = k[len("final."):]
        if name.endswith("num_batches_tracked"):
            assert int(sd[name]) == int(v), name
            continue
        if sage and bn and name.endswith("fc_neigh.bias") and f"layers.{L - 1}." not in name:
            continue
        if sage and bn and name.endswith("running_mean"):
            assert relerr(sd[name].cpu(), v) < 5e-2, name     # carries the noise-driven bias
            continue
        assert relerr_q(sd[name].cpu(), v, 0.99) < q_tol, name


@pytest.mark.parametrize("case", ["teacher_train_gcn3_bn", "teacher_train_gcn2_drop"])
def test_gcn_train_with_batchnorm_and_dropout_matches_reference(dev, case):
    """`train` (train_and_eval.py:12-29) with BatchNorm in train mode / with the dropout keep-masks
    the reference drew, against the reference's own losses and final parameters."""
    from glnn_b200 import graph as G, teacher_train as TT
    from glnn_b200.models import Model
    d = load(case)
    L = int(d["num_layers"])
    model = Model(dict(model_name="GCN", num_layers=L, feat_dim=d["feats"].shape[1],
                       hidden_dim=int(d["hidden"]), label_dim=d["out"].shape[1],
                       dropout_ratio=float(d["dropout"]), norm_type=str(d["norm"]), device=dev))
    model.load_state_dict({k[len("init."):]: torch.from_numpy(np.array(v)) for k, v in d.items()
                           if k.startswith("init.")})
    g = G.graph((d["src"], d["dst"]), num_nodes=int(d["n"])).to(dev)
    feats, labels = torch.from_numpy(d["feats"]).to(dev), torch.from_numpy(d["labels"]).to(dev)
    idx_train = torch.from_numpy(d["idx_train"]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=float(d["lr"]), weight_decay=float(d["wd"]))
    masks = [m.to(dev) for m in _masks_of(d)]
    per_step = L - 1
    losses = []
    for s in range(len(d["losses"])):
        km = masks[s * per_step:(s + 1) * per_step] if masks else None
        losses.append(TT.gcn_train_step(model, g, feats, labels, torch.nn.NLLLoss(), opt, idx_train,
                                        float(d["lamb"]), keep_masks=km).item())
    assert np.allclose(losses, d["losses"], rtol=1e-4), (losses, d["losses"])
    _check_final(model, d)


@pytest.mark.parametrize("case", ["teacher_train_sage3_bn", "teacher_train_sage2_bn_drop"])
def test_sage_train_with_batchnorm_and_dropout_matches_reference(dev, case):
    """`train_sage` (train_and_eval.py:32-56) over full-neighbour blocks built by the DEVICE sampler,
    with BatchNorm in train mode / recorded dropout keep-masks, against the reference's run."""
    from glnn_b200 import graph as G, teacher_train as TT
    from glnn_b200.models import Model
    d = load(case)
    L = int(d["num_layers"])
    init = {k[len("init."):]: torch.from_numpy(np.array(v)) for k, v in d.items() if k.startswith("init.")}
    model = Model(dict(model_name="SAGE", num_layers=L, feat_dim=d["feats"].shape[1],
                       hidden_dim=int(d["hidden"]),
                       label_dim=init["encoder.layers.%d.fc_neigh.weight" % (L - 1)].shape[0],
                       dropout_ratio=float(d["dropout"]), norm_type=str(d["norm"]), device=dev))
    model.load_state_dict(init)
    g = G.graph((d["src"], d["dst"]), num_nodes=int(d["n"])).to(dev)
    feats, labels = torch.from_numpy(d["feats"]).to(dev), torch.from_numpy(d["labels"]).to(dev)
    seeds = torch.from_numpy(d["seeds"]).to(dev)
    blocks, cur = [], seeds
    for _ in range(L):
        cur, blk = TT.sample_block(g, cur, -1)
        blocks.insert(0, blk)
    # the reference's blocks list the extra src nodes in np.setdiff1d (increasing id) order too, so the
    # recorded [n_dst, hidden] masks line up row by row
    opt = torch.optim.Adam(model.parameters(), lr=float(d["lr"]), weight_decay=float(d["wd"]))
    masks = [m.to(dev) for m in _masks_of(d)]
    per_step = L - 1
    model.train()
    losses = []
    for s in range(len(d["losses"])):
        km = masks[s * per_step:(s + 1) * per_step] if masks else None
        losses.append(TT.sage_train_step(model, blocks, feats[cur], labels, seeds, torch.nn.NLLLoss(), opt,
                                         float(d["lamb"]), keep_masks=km).item())
    assert np.allclose(losses, d["losses"], rtol=1e-4), (losses, d["losses"])
    _check_final(model, d, sage=True)
