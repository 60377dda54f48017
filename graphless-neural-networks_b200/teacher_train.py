"""Teacher TRAINING on the B200 kernels -- the first 'next' row of the scope table (SURVEY.md
section 8f): full-batch GCN steps (`train`, train_and_eval.py:12-29) and sampled-block GraphSAGE
steps (`train_sage`, :32-56).  The forward and backward aggregations and projections run on
glnn_spmm_csr_f32 / glnn_gemm_f32 through two small autograd Functions (the backward of a CSR
aggregation is the aggregation over the transposed CSR with the same scale vectors); normalisation,
dropout and the loss stay torch ops here.  Neighbour sampling (dgl MultiLayerNeighborSampler +
NodeDataLoader, train_and_eval.py:179-190) is done on the device with torch primitives.
"""
import torch
import torch.nn.functional as F

from . import ops
from .graph import CSRGraph


# ------------------------------------------------------------------------------------------------
# autograd wrappers over the kernels
# ------------------------------------------------------------------------------------------------
class _Aggregate(torch.autograd.Function):
    """y = dst_scale * (A x [+ x[:n_dst]]) with A given as CSR over destinations; the transposed
    CSR (rows = sources) drives the backward."""

    @staticmethod
    def forward(ctx, x, fwd, bwd, dst_scale, self_add):
        indptr, indices = fwd
        y = ops.spmm_csr(indptr, indices, x.contiguous(), self_add=self_add, dst_scale=dst_scale)
        ctx.bwd, ctx.self_add, ctx.n_src = bwd, self_add, x.shape[0]
        ctx.save_for_backward(dst_scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        (dst_scale,) = ctx.saved_tensors
        t_indptr, t_indices = ctx.bwd
        dy = dy.contiguous()
        # dx = A^T (dst_scale * dy): scale the gathered rows by their (destination) scale
        dx = ops.spmm_csr(t_indptr, t_indices, dy, src_scale=dst_scale)
        if ctx.self_add:
            n_dst = dy.shape[0]
            g = dy if dst_scale is None else dy * dst_scale.unsqueeze(1)
            dx[:n_dst] += g
        return dx, None, None, None, None


class _Linear(torch.autograd.Function):
    """y = x @ op(W) + b on glnn_gemm_f32; weight_is_out_in = nn.Linear layout [out, in]."""

    @staticmethod
    def forward(ctx, x, w, b, weight_is_out_in):
        x = x.contiguous()
        ctx.save_for_backward(x, w)
        ctx.oi, ctx.has_bias = weight_is_out_in, b is not None
        return ops.gemm(x, w, trans_b=weight_is_out_in, bias=b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.oi:  # y = x W^T : dx = dy W, dW = dy^T x
            dx = ops.gemm(dy, w)
            dw = ops.gemm(dy, x, trans_a=True)
        else:       # y = x W   : dx = dy W^T, dW = x^T dy
            dx = ops.gemm(dy, w, trans_b=True)
            dw = ops.gemm(x, dy, trans_a=True)
        return dx, dw, (dy.sum(0) if ctx.has_bias else None), None


def _transpose_csr(indptr, indices, n_src):
    """CSR over sources of the same edge set (for the backward aggregation)."""
    n_dst = indptr.numel() - 1
    deg = (indptr[1:] - indptr[:-1]).to(torch.int64)
    dst = torch.repeat_interleave(torch.arange(n_dst, device=indices.device), deg)
    src = indices.to(torch.int64)
    order = torch.sort(src, stable=True).indices
    t_indices = dst[order].to(torch.int32)
    t_indptr = torch.zeros(n_src + 1, dtype=torch.int64, device=indices.device)
    torch.cumsum(torch.bincount(src, minlength=n_src), 0, out=t_indptr[1:])
    return t_indptr.to(indptr.dtype), t_indices


def _graph_pair(g):
    if getattr(g, "_t_csr", None) is None:
        g._t_csr = _transpose_csr(g.indptr, g.indices, g.num_nodes())
    return (g.indptr, g.indices), g._t_csr


# ------------------------------------------------------------------------------------------------
# GCN full-batch training forward (models.py:189-199 in train mode)
# ------------------------------------------------------------------------------------------------
def gcn_forward_train(enc, g, feats):
    if not isinstance(g, CSRGraph):
        raise TypeError("expected a glnn_b200 CSRGraph")
    g = g.to(feats.device)
    if g.has_zero_in_degree():
        raise ValueError("There are 0-in-degree nodes in the graph (dgl GraphConv would raise)")
    fwd, bwd = _graph_pair(g)
    ns, nd = g.gcn_norms()
    h, h_list = feats, []
    for l, conv in enumerate(enc.layers):
        d_in, d_out = conv.weight.shape
        h = h * ns.unsqueeze(1)
        if d_in > d_out:
            h = _Linear.apply(h, conv.weight, None, False)
            h = _Aggregate.apply(h, fwd, bwd, None, False)
        else:
            h = _Aggregate.apply(h, fwd, bwd, None, False)
            h = _Linear.apply(h, conv.weight, None, False)
        h = h * nd.unsqueeze(1) + conv.bias
        if conv._activation is not None:
            h = conv._activation(h)
        if l != enc.num_layers - 1:
            h_list.append(h)
            if enc.norm_type != "none":
                h = enc.norms[l](h)
            h = enc.dropout(h)
    return h_list, h


# ------------------------------------------------------------------------------------------------
# neighbour sampling + SAGE block training (train_and_eval.py:32-56, 179-190)
# ------------------------------------------------------------------------------------------------
class Block:
    """Bipartite message-flow graph of one layer: CSR over the dst nodes with LOCAL src ids; the dst
    nodes are the first n_dst entries of the src node list (the invariant models.py:105-109 uses)."""

    def __init__(self, indptr, indices, n_src, n_dst):
        self.indptr, self.indices, self.n_src, self.n_dst = indptr, indices, n_src, n_dst
        self._t = None

    def num_dst_nodes(self):
        return self.n_dst

    def num_src_nodes(self):
        return self.n_src

    def int(self):
        return self

    def to(self, device):
        return self

    def pair(self):
        if self._t is None:
            self._t = _transpose_csr(self.indptr, self.indices, self.n_src)
        return (self.indptr, self.indices), self._t


def sample_block(g, seeds, fanout, gen=None):
    """Uniform sampling without replacement of at most `fanout` in-edges per seed."""
    dev = g.indices.device
    p = g.indptr.to(torch.int64)
    start, deg = p[seeds], p[seeds + 1] - p[seeds]
    total = int(deg.sum())
    row = torch.repeat_interleave(torch.arange(seeds.numel(), device=dev), deg)
    first = torch.cumsum(deg, 0) - deg
    pos = start[row] + (torch.arange(total, device=dev) - first[row])
    if fanout is not None and fanout >= 0 and total > 0 and int(deg.max()) > fanout:
        key = torch.rand(total, device=dev, generator=gen)
        order = torch.sort(row.to(torch.float64) + key.to(torch.float64) * 0.999999).indices
        rank = torch.arange(total, device=dev) - first[row[order]]
        keep = order[rank < fanout]
        keep = keep.sort().values
        row, pos = row[keep], pos[keep]
    nbr = g.indices[pos].to(torch.int64)
    # src node list: seeds first, then the other sampled nodes
    flag = torch.zeros(g.num_nodes(), dtype=torch.bool, device=dev)
    flag[nbr] = True
    flag[seeds] = False
    extra = flag.nonzero(as_tuple=True)[0]
    src_nodes = torch.cat([seeds, extra])
    local = torch.empty(g.num_nodes(), dtype=torch.int64, device=dev)
    local[src_nodes] = torch.arange(src_nodes.numel(), device=dev)
    counts = torch.bincount(row, minlength=seeds.numel())
    indptr = torch.zeros(seeds.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    block = Block(indptr.to(torch.int32), local[nbr].to(torch.int32), src_nodes.numel(), seeds.numel())
    return src_nodes, block


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
        for i in range(len(self)):
            seeds = nids[i * self.batch_size:(i + 1) * self.batch_size]
            out_nodes, blocks, cur = seeds, [], seeds
            for fanout in reversed(self.fanouts):
                cur, blk = sample_block(self.g, cur, fanout)
                blocks.insert(0, blk)
            yield cur, out_nodes, blocks


def sage_forward_blocks(enc, blocks, feats):
    """SAGE.forward over sampled blocks in train mode (models.py:101-119)."""
    h, h_list = feats, []
    for l, (conv, blk) in enumerate(zip(enc.layers, blocks)):
        fwd, bwd = blk.pair()
        deg = (blk.indptr[1:] - blk.indptr[:-1]).to(torch.float32)
        inv = 1.0 / (deg + 1.0)
        agg = _Aggregate.apply(h, fwd, bwd, inv, True)
        h = _Linear.apply(agg, conv.fc_neigh.weight, conv.fc_neigh.bias, True)
        if l != enc.num_layers - 1:
            h_list.append(h)
            if enc.norm_type != "none":
                h = enc.norms[l](h)
            h = enc.dropout(enc.activation(h))
    return h_list, h


def train_sage(model, dataloader, feats, labels, criterion, optimizer, lamb=1):
    model.train()
    total_loss, steps = 0.0, 0
    for input_nodes, output_nodes, blocks in dataloader:
        out = model(blocks, feats[input_nodes]).log_softmax(dim=1)
        loss = criterion(out, labels[output_nodes])
        total_loss += loss.item()
        optimizer.zero_grad()
        (loss * lamb).backward()
        optimizer.step()
        steps += 1
    return total_loss / max(steps, 1)
