"""Teacher TRAINING on the B200 kernels -- 'next' rows 1 and 2 of the scope table (SURVEY.md section
8f): full-batch GCN steps (`train`, train_and_eval.py:12-29) and sampled-block GraphSAGE steps
(`train_sage`, :32-56) as explicit kernel sequences with a HAND-WRITTEN backward.  There is no autograd
on this path: every step is

    forward   aggregation (glnn_spmm_csr_f32) / projection (glnn_gemm_f32) with the bias, degree
              scalings and GraphConv's ReLU fused into their epilogues, then the train-mode block
              [BatchNorm1d] -> [ReLU] -> [Dropout] (glnn_act_train_fwd_f32)
    loss      log_softmax + NLL + d(lamb * loss)/dlogits in one kernel (glnn_nll_loss_grad_f32)
    backward  the block's backward incl. the bias gradient (glnn_act_train_bwd_f32), the two
              projection gradients (glnn_gemm_f32) and the transposed aggregation -- over the
              transposed CSR for the fixed full-batch graph, by scatter for sampled blocks
              (glnn_spmm_csr_scatter_f32)
    update    torch.optim.Adam's rule by glnn_adam_step_f32, on the optimizer's own state tensors

Neighbour sampling (dgl MultiLayerNeighborSampler + NodeDataLoader, train_and_eval.py:179-190) runs
on the device too: glnn_sample_neighbors draws the edges, glnn_block_mark / glnn_block_relabel build
the block's local ids."""
import torch

from . import ops
from .graph import CSRGraph


def _transpose_csr(indptr, indices, n_src):
    """CSR over sources of the same edge set (drives the backward aggregation of a fixed graph):
    the edge list in CSR order, re-sorted by source on the device (glnn_csr_from_coo, stable)."""
    n_dst = indptr.numel() - 1
    deg = (indptr[1:] - indptr[:-1]).to(torch.int64)
    dst = torch.repeat_interleave(torch.arange(n_dst, device=indices.device, dtype=torch.int32), deg)
    t_indptr, t_indices, _ = ops.csr_from_coo(dst, indices, n_src, want_out_deg=False)
    return t_indptr.to(indptr.dtype), t_indices


def _graph_pair(g):
    if getattr(g, "_t_csr", None) is None:
        g._t_csr = _transpose_csr(g.indptr, g.indices, g.num_nodes())
    return (g.indptr, g.indices), g._t_csr


# ------------------------------------------------------------------------------------------------
# optimizer: torch.optim.Adam's update by glnn_adam_step_f32 on the optimizer's own state
# ------------------------------------------------------------------------------------------------
def _supported_adam(optimizer):
    if type(optimizer) is not torch.optim.Adam:
        return False
    return all(not g.get("amsgrad") and not g.get("maximize") for g in optimizer.param_groups)


def adam_apply(optimizer, grads):
    """One optimizer step from explicit gradients {parameter: grad tensor}; state["step"],
    ["exp_avg"], ["exp_avg_sq"] are created and advanced exactly like torch.optim.Adam.step()."""
    for group in optimizer.param_groups:
        for p in group["params"]:
            g = grads.get(p)
            if g is None:
                continue
            st = optimizer.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] += 1
            ops.adam_step(p.data.view(-1), g.contiguous().view(-1), st["exp_avg"].view(-1),
                          st["exp_avg_sq"].view(-1), int(st["step"]), group["lr"], group["betas"],
                          group["eps"], group["weight_decay"])


def _dropout_seed():
    """A fresh 62-bit seed per step from torch's CPU generator (so set_seed() makes runs repeatable)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _nll_only(criterion):
    import torch.nn as nn
    if not (type(criterion) is nn.NLLLoss and criterion.reduction == "mean" and criterion.weight is None
            and criterion.ignore_index == -100):
        raise NotImplementedError("teacher training kernels implement the reference's criterion, "
                                  "torch.nn.NLLLoss() (train_teacher.py:237)")


def _act_block(x, enc, l, relu_post, relu_input, keep_masks):
    """The train-mode block after hidden layer l."""
    if enc.norm_type == "layer":
        raise NotImplementedError("norm_type='layer' is outside the B200 hot path")
    bn = enc.norms[l] if enc.norm_type == "batch" else None
    p_drop = float(enc.dropout.p)
    mask = None if keep_masks is None else keep_masks[l]
    if mask is not None:
        return ops.ActBlock(x, bn=bn, relu_post=relu_post, relu_input=relu_input, p_drop=p_drop,
                            keep_mask=mask)
    return ops.ActBlock(x, bn=bn, relu_post=relu_post, relu_input=relu_input, p_drop=p_drop,
                        seed=_dropout_seed() if p_drop > 0 else 0)


# ------------------------------------------------------------------------------------------------
# GCN: full-batch step (models.py:189-199 in train mode + train_and_eval.py:12-29)
# ------------------------------------------------------------------------------------------------
def gcn_forward_train(enc, g, feats, keep_masks=None):
    """GCN.forward in TRAIN mode (models.py:189-199) on the kernels: returns (h_list, logits, cache);
    h_list holds the post-activation conv outputs like the reference, cache what the backward needs."""
    if not isinstance(g, CSRGraph):
        raise TypeError("expected a glnn_b200 CSRGraph")
    g = g.to(feats.device)
    if g.has_zero_in_degree():
        raise ValueError("There are 0-in-degree nodes in the graph (dgl GraphConv would raise)")
    indptr, indices = g.indptr, g.indices
    ns, nd = g.gcn_norms()
    L = enc.num_layers
    h = feats.contiguous()
    cache, h_list = [], []
    for l, conv in enumerate(enc.layers):
        d_in, d_out = conv.weight.shape
        relu = 2 if conv._activation is not None else 0
        if d_in > d_out:   # DGL: project first when it shrinks the rows
            z = ops.gemm(h, conv.weight, row_scale=ns)
            y = ops.spmm_csr(indptr, indices, z, dst_scale=nd, bias=conv.bias, relu=relu)
            saved = ("project_first", h)
        else:
            t = ops.spmm_csr(indptr, indices, h, src_scale=ns)
            y = ops.gemm(t, conv.weight, row_scale=nd, bias=conv.bias, relu=relu)
            saved = ("aggregate_first", t)
        blk = None
        if l != L - 1:
            blk = _act_block(y, enc, l, False, relu != 0, keep_masks)
            h = blk.forward()
        else:
            h = y
        if l != L - 1:
            h_list.append(y)
        cache.append((saved, y, relu, blk))
    return h_list, h, cache


def gcn_train_step(model, g, feats, labels, criterion, optimizer, idx_train, lamb=1, keep_masks=None):
    """One `train` step.  keep_masks: optional list of uint8 [n, hidden] dropout keep-masks, one per
    hidden layer (parity mode); otherwise a counter-based device stream.  Returns the device scalar
    holding the unscaled loss."""
    _nll_only(criterion)
    if not _supported_adam(optimizer):
        raise NotImplementedError("teacher training kernels implement torch.optim.Adam "
                                  "(train_teacher.py:234-236)")
    enc = model.encoder
    model.train()
    g = g.to(feats.device) if isinstance(g, CSRGraph) else g
    _, h, cache = gcn_forward_train(enc, g, feats, keep_masks)
    (indptr, indices), (t_indptr, t_indices) = _graph_pair(g)
    ns, nd = g.gcn_norms()
    L = enc.num_layers
    idx_train = torch.as_tensor(idx_train, dtype=torch.int64, device=feats.device)
    dlog, loss = ops.nll_loss_grad(h, labels, rows=idx_train, lamb=lamb)
    grads = {}
    d = dlog
    for l in reversed(range(L)):
        conv = enc.layers[l]
        (order, x_saved), y, relu, blk = cache[l]
        if blk is None:   # last layer: only the conv's own ReLU (1-layer GCN) and the bias gradient
            blk = ops.ActBlock(y, relu_input=relu != 0)
        dz, dgam, dbet, dbias = blk.backward(d)
        if dgam is not None:
            grads[enc.norms[l].weight], grads[enc.norms[l].bias] = dgam, dbet
        grads[conv.bias] = dbias
        if order == "project_first":
            # y = act(nd * A ((ns * h) W) + b):  dm = ns * A^T (nd * dz);  dW = h^T dm;  dh = dm W^T
            dm = ops.spmm_csr(t_indptr, t_indices, dz, src_scale=nd, dst_scale=ns)
            grads[conv.weight] = ops.gemm(x_saved, dm, trans_a=True)
            if l > 0:
                d = ops.gemm(dm, conv.weight, trans_b=True)
        else:
            # y = act(nd * (A (ns * h)) W + b):  dW = t^T (nd * dz);  dh = ns * A^T ((nd * dz) W^T)
            dzs = dz * nd.unsqueeze(1)
            grads[conv.weight] = ops.gemm(x_saved, dzs, trans_a=True)
            if l > 0:
                dt = ops.gemm(dzs, conv.weight, trans_b=True)
                d = ops.spmm_csr(t_indptr, t_indices, dt, dst_scale=ns)
    adam_apply(optimizer, grads)
    return loss


# ------------------------------------------------------------------------------------------------
# neighbour sampling + SAGE block training (train_and_eval.py:32-56, 179-190)
# ------------------------------------------------------------------------------------------------
class Block:
    """Bipartite message-flow graph of one layer: CSR over the dst nodes with LOCAL src ids; the dst
    nodes are the first n_dst entries of the src node list (the invariant models.py:105-109 uses)."""

    def __init__(self, indptr, indices, n_src, n_dst):
        self.indptr, self.indices, self.n_src, self.n_dst = indptr, indices, n_src, n_dst
        deg = (indptr[1:] - indptr[:-1]).to(torch.float32)
        self.inv_deg1 = (1.0 / (deg + 1.0)).contiguous()

    def num_dst_nodes(self):
        return self.n_dst

    def num_src_nodes(self):
        return self.n_src

    def int(self):
        return self

    def to(self, device):
        return self


def sample_block(g, seeds, fanout, rng_seed=0):
    """One hop: at most `fanout` in-edges per seed, uniformly without replacement (fanout None or < 0:
    all of them), as a Block over local ids.  Returns (src_nodes int64, block); the seeds are the
    first len(seeds) entries of src_nodes, the other sampled nodes follow in increasing id order."""
    dev = g.indices.device
    seeds = seeds.to(torch.int64)
    m = seeds.numel()
    fan = -1 if fanout is None else int(fanout)
    blk_ptr, src = ops.sample_neighbors(g.indptr, g.indices, seeds, fan, rng_seed)
    # local ids: seeds first, then every other touched node
    flag = torch.zeros(g.num_nodes(), dtype=torch.uint8, device=dev)
    ops.block_mark(src, flag)
    flag[seeds] = 0
    extra = flag.nonzero(as_tuple=True)[0]
    src_nodes = torch.cat([seeds, extra])
    node_map = torch.empty(g.num_nodes(), dtype=torch.int32, device=dev)
    node_map[src_nodes] = torch.arange(src_nodes.numel(), dtype=torch.int32, device=dev)
    ops.block_relabel(src, node_map)
    indptr = blk_ptr.to(torch.int32) if src.numel() < 2 ** 31 else blk_ptr
    return src_nodes, Block(indptr, src, src_nodes.numel(), m)


class NeighborLoader:
    """NodeDataLoader(g, nids, MultiLayerNeighborSampler(fan_out), batch_size, shuffle) look-alike:
    yields (input_nodes, output_nodes, blocks) with blocks[0] the outermost hop.  `fan_out` is the
    reference's comma-separated string (train.conf.yaml) or a list."""

    def __init__(self, g, nids, fan_out, batch_size, shuffle=True, drop_last=False):
        self.g = g
        self.nids = torch.as_tensor(nids, dtype=torch.int64)
        self.fanouts = [int(x) for x in fan_out.split(",")] if isinstance(fan_out, str) \
            else list(fan_out)
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last

    def __len__(self):
        n = self.nids.numel()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        dev = self.g.indices.device
        nids = self.nids[torch.randperm(self.nids.numel())] if self.shuffle else self.nids
        nids = nids.to(dev)
        base = int(torch.randint(0, 2 ** 62, (1,)).item())   # CPU generator: repeatable under set_seed
        for i in range(len(self)):
            seeds = nids[i * self.batch_size:(i + 1) * self.batch_size]
            out_nodes, blocks, cur = seeds, [], seeds
            for hop, fanout in enumerate(reversed(self.fanouts)):
                cur, blk = sample_block(self.g, cur, fanout, base + 1000003 * i + hop)
                blocks.insert(0, blk)
            yield cur, out_nodes, blocks


def sage_forward_blocks(enc, blocks, feats_in, keep_masks=None):
    """SAGE.forward over sampled blocks in TRAIN mode (models.py:101-119) on the kernels: returns
    (h_list, logits, cache); h_list holds the conv outputs before norm / activation like the
    reference."""
    L = enc.num_layers
    h = feats_in.contiguous()
    cache, h_list = [], []
    for l, (conv, blk) in enumerate(zip(enc.layers, blocks)):
        agg = ops.spmm_csr(blk.indptr, blk.indices, h, self_add=True, dst_scale=blk.inv_deg1)
        z = ops.gemm(agg, conv.fc_neigh.weight, trans_b=True, bias=conv.fc_neigh.bias)
        act = None
        if l != L - 1:
            act = _act_block(z, enc, l, True, False, keep_masks)
            h = act.forward()
        else:
            h = z
        if l != L - 1:
            h_list.append(z)
        cache.append((agg, z, act))
    return h_list, h, cache


def sage_train_step(model, blocks, feats_in, labels, label_rows, criterion, optimizer, lamb=1,
                    keep_masks=None):
    """One optimizer step of `train_sage` on one batch of blocks.  feats_in: features of the input
    nodes [n_src0, F]; the labels of the seeds are labels[label_rows].  Returns the device scalar
    holding the unscaled loss."""
    _nll_only(criterion)
    if not _supported_adam(optimizer):
        raise NotImplementedError("teacher training kernels implement torch.optim.Adam")
    enc = model.encoder
    L = enc.num_layers
    _, h, cache = sage_forward_blocks(enc, blocks, feats_in, keep_masks)
    dlog, loss = ops.nll_loss_grad(h, labels, label_rows=label_rows, lamb=lamb)
    grads = {}
    d = dlog
    for l in reversed(range(L)):
        conv, blk = enc.layers[l], blocks[l]
        agg, z, act = cache[l]
        if act is None:
            act = ops.ActBlock(z)            # identity block: only the bias gradient (column sums)
        dz, dgam, dbet, dbias = act.backward(d)
        if dgam is not None:
            grads[enc.norms[l].weight], grads[enc.norms[l].bias] = dgam, dbet
        grads[conv.fc_neigh.bias] = dbias
        grads[conv.fc_neigh.weight] = ops.gemm(dz, agg, trans_a=True)
        if l > 0:
            dagg = ops.gemm(dz, conv.fc_neigh.weight)
            d = torch.zeros(blk.n_src, agg.shape[1], dtype=torch.float32, device=dz.device)
            ops.spmm_scatter(blk.indptr, blk.indices, dagg, blk.inv_deg1, d, self_add=True)
    adam_apply(optimizer, grads)
    return loss


def train_sage(model, dataloader, feats, labels, criterion, optimizer, lamb=1):
    """train_and_eval.py:32-56.  The per-step losses stay on the device; ONE host read per pass (the
    reference reads every step, :47)."""
    model.train()
    total, steps = None, 0
    for input_nodes, output_nodes, blocks in dataloader:
        input_nodes = torch.as_tensor(input_nodes, device=feats.device)
        output_nodes = torch.as_tensor(output_nodes, dtype=torch.int64, device=feats.device)
        loss = sage_train_step(model, blocks, feats[input_nodes], labels, output_nodes, criterion,
                               optimizer, lamb)
        total = loss if total is None else total + loss
        steps += 1
    return float(total.item()) / steps if steps else 0.0
