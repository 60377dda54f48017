"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal stand-in for the third-party libraries the reference imports but which are not
installable here (dgl==0.6.1, ogb==1.3.3, pytz -- /root/reference/requirements.txt:3-14), so that
/root/reference/{models,train_and_eval,utils}.py can be imported and executed UNMODIFIED on CPU in
the authoring container.  It is used by oracle/make_golden.py to produce tests/golden/*.npz.

Every piece of DGL 0.6.1 semantics that the hot path depends on is restated in this one file, one
rule per comment, so that a disagreement with a real DGL can be fixed in a single place
(SURVEY.md section 8c "risk").  The rules come from the published DGL 0.6.x sources
(python/dgl/nn/pytorch/conv/sageconv.py, graphconv.py, python/dgl/dataloading/*):

 R1  SAGEConv(in, out, "gcn"): neigh = sum_{u->v} h_u (multi-edges counted with multiplicity);
     h_neigh = (neigh + h_dst) / (in_deg(v) + 1); rst = fc_neigh(h_neigh).  There is no fc_self for
     the "gcn" aggregator and 0.6.1 always aggregates before projecting.
 R2  SAGEConv.reset_parameters: xavier_uniform_(fc_neigh.weight, gain=calculate_gain("relu")); the
     bias keeps nn.Linear's default init.
 R3  GraphConv(in, out, norm="both", weight=True, bias=True, activation): feat * out_deg^-1/2 (clamp
     min 1); if in > out: (feat @ W) then sum-aggregate, else aggregate then @ W; * in_deg^-1/2
     (clamp min 1); + bias; activation.  W has shape [in, out] (xavier_uniform_, gain 1), bias zeros.
     A graph with a zero-in-degree node raises DGLError unless allow_zero_in_degree.
 R4  MultiLayerFullNeighborSampler(1) + NodeDataLoader(g, nids, sampler, batch_size, shuffle=False,
     drop_last=False): yields (input_nodes, output_nodes, [block]); block holds every in-edge of the
     batch's dst nodes; input_nodes starts with the dst nodes (the invariant
     /root/reference/models.py:105-109,137 relies on).
 R5  block.int()/.to(device)/graph.to(device)/create_formats_() do not change the structure.
"""
import sys
import types

import numpy as np
import torch
import torch.nn as nn


class DGLError(Exception):
    pass


class ShimGraph:
    """Directed multigraph stored as CSR over destination nodes (row v lists the sources of v's
    in-edges), which is all the hot path needs."""

    def __init__(self, src, dst, num_nodes):
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        self._n = int(num_nodes)
        order = np.argsort(dst, kind="stable")
        self.indices = torch.from_numpy(src[order].copy())
        counts = np.bincount(dst, minlength=self._n)
        indptr = np.zeros(self._n + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        self.indptr = torch.from_numpy(indptr)
        self._out_deg = torch.from_numpy(np.bincount(src, minlength=self._n).astype(np.int64))
        self.ndata = {}

    # -- the slice of the DGLGraph API the reference touches on the hot path --------------------
    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self.indices.numel())

    def in_degrees(self):
        return self.indptr[1:] - self.indptr[:-1]

    def out_degrees(self):
        return self._out_deg

    def create_formats_(self):  # R5
        return None

    def to(self, device):  # R5
        return self

    def int(self):  # R5
        return self

    def spmm_sum(self, h):
        """copy_u -> sum: out[v] = sum over in-edges (u->v) of h[u]."""
        vals = torch.ones(self.indices.numel(), dtype=h.dtype)
        a = torch.sparse_csr_tensor(self.indptr, self.indices, vals, size=(self._n, self._n))
        return a @ h


class Block:
    """Bipartite message-flow graph for one batch (R4)."""

    def __init__(self, indptr, indices, n_src, n_dst):
        self.indptr, self.indices, self.n_src, self.n_dst = indptr, indices, n_src, n_dst

    def num_dst_nodes(self):
        return self.n_dst

    def num_src_nodes(self):
        return self.n_src

    def in_degrees(self):
        return self.indptr[1:] - self.indptr[:-1]

    def int(self):  # R5
        return self

    def to(self, device):  # R5
        return self

    def spmm_sum(self, h_src):
        vals = torch.ones(self.indices.numel(), dtype=h_src.dtype)
        a = torch.sparse_csr_tensor(self.indptr, self.indices, vals, size=(self.n_dst, self.n_src))
        return a @ h_src


class SAGEConv(nn.Module):
    def __init__(self, in_feats, out_feats, aggregator_type, bias=True):
        super().__init__()
        if aggregator_type != "gcn":
            raise NotImplementedError("shim only restates the 'gcn' aggregator (R1)")
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=bias)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=nn.init.calculate_gain("relu"))  # R2

    def forward(self, graph, feat):
        feat_src, feat_dst = feat if isinstance(feat, tuple) else (feat, feat)
        neigh = graph.spmm_sum(feat_src)  # R1
        degs = graph.in_degrees().to(feat_dst.dtype)
        h_neigh = (neigh + feat_dst) / (degs.unsqueeze(-1) + 1)
        return self.fc_neigh(h_neigh)


class GraphConv(nn.Module):
    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        assert norm == "both" and weight and bias
        self._in, self._out = in_feats, out_feats
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.zeros(out_feats))
        nn.init.xavier_uniform_(self.weight)  # R3
        self._activation = activation
        self._allow_zero = allow_zero_in_degree

    def forward(self, graph, feat):
        if not self._allow_zero and bool((graph.in_degrees() == 0).any()):
            raise DGLError("There are 0-in-degree nodes in the graph")
        norm_src = graph.out_degrees().to(feat.dtype).clamp(min=1).pow(-0.5).unsqueeze(-1)
        h = feat * norm_src
        if self._in > self._out:  # R3: project first when it shrinks the rows
            rst = graph.spmm_sum(h @ self.weight)
        else:
            rst = graph.spmm_sum(h) @ self.weight
        norm_dst = graph.in_degrees().to(feat.dtype).clamp(min=1).pow(-0.5).unsqueeze(-1)
        rst = rst * norm_dst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class _Unsupported(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("GATConv/APPNPConv are out of scope (SURVEY.md section 2.1 #5)")


class MultiLayerFullNeighborSampler:
    def __init__(self, n_layers):
        assert n_layers == 1
        self.n_layers = n_layers


class MultiLayerNeighborSampler:
    def __init__(self, fanouts):
        self.fanouts = fanouts


class NodeDataLoader:
    """R4.  Only the full-neighbour, unshuffled form used by SAGE.inference is restated."""

    def __init__(self, g, nids, sampler, batch_size=1, shuffle=False, drop_last=False, num_workers=0):
        self.g, self.sampler = g, sampler
        self.nids = torch.as_tensor(nids, dtype=torch.int64)
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last

    def __len__(self):
        n = self.nids.numel()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        if not isinstance(self.sampler, MultiLayerFullNeighborSampler):
            raise NotImplementedError("fan-out sampling is a next row (SURVEY.md section 8f)")
        nids = self.nids[torch.randperm(self.nids.numel())] if self.shuffle else self.nids
        indptr, indices = self.g.indptr.numpy(), self.g.indices.numpy()
        for i in range(len(self)):
            out_nodes = nids[i * self.batch_size:(i + 1) * self.batch_size]
            yield make_block(indptr, indices, out_nodes.numpy())


def make_block(indptr, indices, out_nodes):
    """All in-edges of `out_nodes`, sources relabelled so that the dst nodes are the prefix."""
    starts, ends = indptr[out_nodes], indptr[out_nodes + 1]
    lens = ends - starts
    b_indptr = np.zeros(len(out_nodes) + 1, dtype=np.int64)
    np.cumsum(lens, out=b_indptr[1:])
    # flat positions of every in-edge of the batch
    pos = np.repeat(starts - b_indptr[:-1], lens) + np.arange(b_indptr[-1])
    src = indices[pos]
    extra = np.setdiff1d(src, out_nodes)  # sorted unique sources that are not dst nodes
    input_nodes = np.concatenate([out_nodes, extra])
    # relabel: position of each global id in input_nodes
    order = np.argsort(input_nodes, kind="stable")
    local = order[np.searchsorted(input_nodes[order], src)]
    block = Block(torch.from_numpy(b_indptr), torch.from_numpy(local.astype(np.int64)),
                  len(input_nodes), len(out_nodes))
    return torch.from_numpy(input_nodes), torch.from_numpy(np.asarray(out_nodes)), [block]


def graph(data, num_nodes=None):
    src, dst = data
    src, dst = np.asarray(src), np.asarray(dst)
    if num_nodes is None:
        num_nodes = int(max(src.max(), dst.max())) + 1
    return ShimGraph(src, dst, num_nodes)


class _Evaluator:
    def __init__(self, name):
        self.name = name


class _TZ:
    def __init__(self, name):
        self.name = name


def install():
    """Register the stand-ins in sys.modules (idempotent)."""
    if "dgl" in sys.modules and getattr(sys.modules["dgl"], "__glnn_shim__", False):
        return
    dgl = types.ModuleType("dgl")
    dgl.__glnn_shim__ = True
    dgl.graph = graph
    dgl.DGLError = DGLError
    dgl_nn = types.ModuleType("dgl.nn")
    dgl_nn.SAGEConv, dgl_nn.GraphConv = SAGEConv, GraphConv
    dgl_nn.GATConv = dgl_nn.APPNPConv = _Unsupported
    dgl_fn = types.ModuleType("dgl.function")
    dgl_dl = types.ModuleType("dgl.dataloading")
    dgl_dl.MultiLayerFullNeighborSampler = MultiLayerFullNeighborSampler
    dgl_dl.MultiLayerNeighborSampler = MultiLayerNeighborSampler
    dgl_dl.NodeDataLoader = NodeDataLoader
    dgl.nn, dgl.function, dgl.dataloading = dgl_nn, dgl_fn, dgl_dl
    ogb = types.ModuleType("ogb")
    ogb_np = types.ModuleType("ogb.nodeproppred")
    ogb_np.Evaluator = _Evaluator
    ogb.nodeproppred = ogb_np
    pytz = types.ModuleType("pytz")
    pytz.timezone = _TZ
    for name, mod in [("dgl", dgl), ("dgl.nn", dgl_nn), ("dgl.function", dgl_fn),
                      ("dgl.dataloading", dgl_dl), ("ogb", ogb), ("ogb.nodeproppred", ogb_np),
                      ("pytz", pytz)]:
        sys.modules[name] = mod
