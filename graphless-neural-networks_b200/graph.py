"""Graph container handed to the teacher kernels: CSR over destination nodes (row v = sources of
v's in-edges, multi-edges kept), int32 ids.  It offers the slice of the DGLGraph API the
reference's scripts touch (dataloader.py:61-111, train_and_eval.py:178-211, utils.py:171-189) so
that those call sites keep working without DGL."""
import numpy as np
import torch


class CSRGraph:
    def __init__(self, indptr, indices, num_nodes, out_deg=None):
        self.indptr = indptr      # int32 (int64 when nnz >= 2^31) [n+1]
        self.indices = indices    # int32 [nnz]
        self._n = int(num_nodes)
        self._out_deg = out_deg   # int64 [n]
        self.ndata = {}
        self._norms = None

    # -- construction -----------------------------------------------------------------------
    @staticmethod
    def from_edges(src, dst, num_nodes=None, device=None):
        """CSR with the edges of every row in their input order (stable by destination).  On a CUDA
        device this is glnn_csr_from_coo (csrc/csr_build.cu: degree counts, scan, 8-bit LSD radix
        passes); edge lists that live on the host are sorted there, as DGL does for the reference."""
        src = torch.as_tensor(src, device=device)
        dst = torch.as_tensor(dst, device=device)
        if src.dtype != torch.int32 or dst.dtype != torch.int32:
            src, dst = src.to(torch.int64), dst.to(torch.int64)
        if num_nodes is None:
            num_nodes = int(max(src.max(), dst.max())) + 1 if src.numel() else 0
        if src.is_cuda:
            from . import ops
            indptr, indices, out_deg = ops.csr_from_coo(src, dst, num_nodes)
            return CSRGraph(indptr, indices, num_nodes, out_deg)
        src, dst = src.to(torch.int64), dst.to(torch.int64)
        order = torch.sort(dst, stable=True).indices
        indices = src[order].to(torch.int32)
        counts = torch.bincount(dst, minlength=num_nodes)
        indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=src.device)
        torch.cumsum(counts, 0, out=indptr[1:])
        if indices.numel() < 2 ** 31:
            indptr = indptr.to(torch.int32)
        out_deg = torch.bincount(src, minlength=num_nodes)
        return CSRGraph(indptr, indices, num_nodes, out_deg)

    # -- DGLGraph look-alikes -----------------------------------------------------------------
    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self.indices.numel())

    number_of_edges = num_edges

    @property
    def device(self):
        return self.indices.device

    def in_degrees(self):
        p = self.indptr.to(torch.int64)
        return p[1:] - p[:-1]

    def out_degrees(self):
        if self._out_deg is None:
            self._out_deg = torch.bincount(self.indices.to(torch.int64), minlength=self._n)
        return self._out_deg

    def create_formats_(self):
        return None

    def int(self):
        return self

    def to(self, device):
        device = torch.device(device)
        if device == self.indices.device:
            return self
        g = CSRGraph(self.indptr.to(device), self.indices.to(device), self._n,
                     None if self._out_deg is None else self._out_deg.to(device))
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        return g

    def edges(self):
        """(src, dst) in CSR order."""
        dst = torch.repeat_interleave(torch.arange(self._n, device=self.device), self.in_degrees())
        return self.indices.to(torch.int64), dst

    def subgraph(self, nodes):
        """Node-induced subgraph with nodes relabelled in the given order (dgl.DGLGraph.subgraph;
        used by the inductive split, train_and_eval.py:324).  The new destination is a function of
        the old one, so on the device no sort is needed: glnn_csr_subgraph compacts every kept row
        in place order."""
        nodes = torch.as_tensor(nodes, dtype=torch.int64, device=self.device)
        if self.indices.is_cuda:
            from . import ops
            relabel = torch.full((self._n,), -1, dtype=torch.int32, device=self.device)
            relabel[nodes] = torch.arange(nodes.numel(), dtype=torch.int32, device=self.device)
            indptr, indices, out_deg = ops.csr_subgraph(self.indptr, self.indices, relabel, nodes.numel())
            g = CSRGraph(indptr, indices, nodes.numel(), out_deg)
        else:
            relabel = torch.full((self._n,), -1, dtype=torch.int64, device=self.device)
            relabel[nodes] = torch.arange(nodes.numel(), device=self.device)
            src, dst = self.edges()
            s, d = relabel[src], relabel[dst]
            keep = (s >= 0) & (d >= 0)
            g = CSRGraph.from_edges(s[keep], d[keep], nodes.numel())
        g.ndata = {k: v[nodes] for k, v in self.ndata.items()}
        return g

    # -- GraphConv norm="both" helpers ----------------------------------------------------------
    def gcn_norms(self):
        """(out_deg^-1/2, in_deg^-1/2), degrees clamped to >= 1 (dgl GraphConv, rule R3)."""
        if self._norms is None:
            ns = self.out_degrees().to(torch.float32).clamp(min=1).pow(-0.5)
            nd = self.in_degrees().to(torch.float32).clamp(min=1).pow(-0.5)
            self._norms = (ns.contiguous(), nd.contiguous())
        return self._norms

    def has_zero_in_degree(self):
        return bool((self.in_degrees() == 0).any())


def graph(data, num_nodes=None, device=None):
    """dgl.graph((src, dst)) look-alike (dataloader.py:78,105)."""
    src, dst = data
    if isinstance(src, np.ndarray):
        src, dst = torch.from_numpy(src), torch.from_numpy(dst)
    return CSRGraph.from_edges(src, dst, num_nodes, device)


class FullNeighborLoader:
    """Stand-in for NodeDataLoader(g, arange(N), MultiLayerFullNeighborSampler(1), ...)
    (train_and_eval.py:193-202): in eval mode the reference's batched layer-wise loop equals one
    full-graph pass per layer, so the loader only has to carry the graph to SAGE.inference."""

    def __init__(self, g, batch_size=None):
        self.g = g
        self.batch_size = batch_size
