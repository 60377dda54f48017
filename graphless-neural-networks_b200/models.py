"""`Model` and its encoders with the reference's surface (/root/reference/models.py:347-429: ctor
keys, substring dispatch order, forward / forward_fitnet / inference, state_dict key names) and the
hot math routed to libglnn_b200.so:

  SAGE.inference (models.py:121-148)  -> glnn_sage_forward   (full-graph layer-wise, eval)
  GCN.forward, eval (models.py:189-199) -> glnn_gcn_forward
  MLP.forward, eval (models.py:42-53)   -> glnn_mlp_eval
  MLP training steps are fused in train_and_eval.train_mini_batch (glnn_mlp_train_pass), which
  never calls forward().

Parameters are ordinary nn.Parameters created by the same torch constructors in the same order as
the reference, so a seeded `Model(conf)` starts from the reference's initial weights.  The MLP
additionally keeps its parameters as views into one flat buffer (see mlp_engine.py).
"""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, mlp_engine, ops
from .graph import CSRGraph, FullNeighborLoader


def _norm_layer(norm_type, dim):
    if norm_type == "batch":
        return nn.BatchNorm1d(dim)
    if norm_type == "layer":
        return nn.LayerNorm(dim)
    return None


def _stack_dims(num_layers, input_dim, hidden_dim, output_dim):
    if num_layers == 1:
        return [(input_dim, output_dim)]
    return [(input_dim, hidden_dim)] + [(hidden_dim, hidden_dim)] * (num_layers - 2) + \
        [(hidden_dim, output_dim)]


class _Encoder(nn.Module):
    """layers / norms ModuleLists laid out as in the reference (a norm after every non-final layer
    when norm_type != "none")."""

    def _build(self, make_layer, num_layers, input_dim, hidden_dim, output_dim, norm_type):
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.norm_type = norm_type
        self.layers = nn.ModuleList()
        self.norms = nn.ModuleList()
        dims = _stack_dims(num_layers, input_dim, hidden_dim, output_dim)
        for i, (d_in, d_out) in enumerate(dims):
            self.layers.append(make_layer(i, d_in, d_out))
            if i != len(dims) - 1:
                nl = _norm_layer(norm_type, d_out)
                if nl is not None:
                    self.norms.append(nl)

    def _gnn_layer_table(self, weights, biases):
        """ctypes glnn_gnn_layer[] + the tensors it points to (kept alive by the caller)."""
        L = self.num_layers
        arr = (_lib.GnnLayer * L)()
        keep = []
        for l in range(L):
            w, b = weights[l], biases[l]
            scale = shift = None
            if l != L - 1 and self.norm_type == "batch":
                bn = self.norms[l]
                scale, shift = ops.bn_fold(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            elif l != L - 1 and self.norm_type == "layer":
                raise NotImplementedError("norm_type='layer' is outside the B200 hot path "
                                          "(no reference config uses it)")
            keep += [w, b, scale, shift]
            arr[l].weight, arr[l].bias = w.data_ptr(), b.data_ptr()
            arr[l].bn_scale = None if scale is None else scale.data_ptr()
            arr[l].bn_shift = None if shift is None else shift.data_ptr()
        return arr, keep

    def _workspace(self, nbytes, device):
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < nbytes or ws.device != device:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws = ws
        return ws


# ------------------------------------------------------------------------------------------------
# MLP (student)
# ------------------------------------------------------------------------------------------------
class MLP(_Encoder):
    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, dropout_ratio,
                 norm_type="none"):
        super().__init__()
        self.input_dim = input_dim
        self.dropout_ratio = dropout_ratio
        self.dropout = nn.Dropout(dropout_ratio)
        self._build(lambda i, a, b: nn.Linear(a, b), num_layers, input_dim, hidden_dim, output_dim,
                    norm_type)
        self._flat = None

    def fused_supported(self):
        return self.norm_type in ("none", "batch")

    def forward(self, feats):
        """(h_list, h) like the reference; h_list holds the pre-norm linear outputs."""
        if feats.is_cuda and not self.training and self.fused_supported() \
                and not torch.is_grad_enabled():
            return [], mlp_engine.eval_logits(self, feats)
        # train-mode / autograd / CPU forward: not a kernel path.  Training goes through
        # train_and_eval.train_mini_batch (fused step); this torch-module forward only runs when the
        # caller opts in (GLNN_ALLOW_TORCH_FALLBACK=1) or asks for forward_fitnet's hidden states.
        from .train_and_eval import torch_fallback_allowed
        if not torch_fallback_allowed():
            raise _lib.GlnnError("MLP.forward outside the fused eval path (train mode, autograd enabled or "
                                 "CPU tensors): use train_mini_batch / evaluate_mini_batch, or set "
                                 "GLNN_ALLOW_TORCH_FALLBACK=1 for the plain torch-module forward")
        return self._forward_autograd(feats)

    def _forward_autograd(self, feats):
        # torch-module path: only for callers that need autograd through forward() or h_list
        # (forward_fitnet); the measured train/eval loops never come here.
        h, h_list = feats, []
        for l, layer in enumerate(self.layers):
            h = layer(h)
            if l != self.num_layers - 1:
                h_list.append(h)
                if self.norm_type != "none":
                    h = self.norms[l](h)
                h = self.dropout(F.relu(h))
        return h_list, h


# ------------------------------------------------------------------------------------------------
# SAGE (teacher)
# ------------------------------------------------------------------------------------------------
class SAGEConv(nn.Module):
    """Parameter holder with dgl.nn.SAGEConv(in, out, "gcn")'s state_dict names and init (rule
    R1/R2 of oracle/dgl_shim.py): fc_neigh only, xavier_uniform with the ReLU gain."""

    def __init__(self, in_feats, out_feats, aggregator_type="gcn"):
        super().__init__()
        if aggregator_type != "gcn":
            raise NotImplementedError("only the 'gcn' aggregator is on the GLNN path")
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=True)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=nn.init.calculate_gain("relu"))


def _fwd_flags(log_softmax, exact):
    import os
    if exact is None:
        exact = os.environ.get("GLNN_EXACT", "0") not in ("", "0")
    return (1 if log_softmax else 0) | (2 if exact else 0)


def _as_graph(data):
    if isinstance(data, FullNeighborLoader):
        return data.g
    if isinstance(data, CSRGraph):
        return data
    raise TypeError("expected a glnn_b200 CSRGraph or FullNeighborLoader, got %r" % type(data))


class SAGE(_Encoder):
    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, dropout_ratio, activation,
                 norm_type="none"):
        super().__init__()
        self.activation = activation
        self.dropout = nn.Dropout(dropout_ratio)
        self._build(lambda i, a, b: SAGEConv(a, b, "gcn"), num_layers, input_dim, hidden_dim,
                    output_dim, norm_type)

    def forward(self, blocks, feats):
        """Sampled-block forward (models.py:101-119) on the kernels, in the module's current mode
        (train: batch statistics + dropout).  It returns plain tensors: there is no autograd through
        glnn_b200 -- training goes through train_and_eval.train_sage, whose backward is a hand-written
        kernel sequence (teacher_train.py)."""
        from . import teacher_train
        if not self.training:
            raise NotImplementedError("block-wise eval forward is not on the GLNN path: the reference "
                                      "evaluates SAGE with inference() (train_and_eval.py:97)")
        h_list, h, _ = teacher_train.sage_forward_blocks(self, blocks, feats)
        return h_list, h

    def inference(self, data, feats, log_softmax=False, exact=None):
        """Full-neighbour layer-wise inference for every node (models.py:121-148).  exact=True (or
        env GLNN_EXACT=1) selects the plain-fp32 mode of glnn_sage_forward (see include/glnn_b200.h)."""
        g = _as_graph(data)
        _lib.require_cuda(feats)
        g = g.to(feats.device)
        lib = _lib.load()
        feats = feats if feats.stride(-1) == 1 else feats.contiguous()
        n = g.num_nodes()
        if feats.shape[0] != n:
            raise ValueError("feats must have one row per graph node")
        layers, keep = self._gnn_layer_table([l.fc_neigh.weight for l in self.layers],
                                             [l.fc_neigh.bias for l in self.layers])
        for l, conv in enumerate(self.layers):
            layers[l].d_out, layers[l].d_in = conv.fc_neigh.weight.shape
        nbytes = lib.glnn_gnn_forward_workspace_bytes(n, layers, self.num_layers)
        ws = self._workspace(nbytes, feats.device)
        out = torch.empty(n, self.layers[-1].fc_neigh.weight.shape[0], dtype=torch.float32,
                          device=feats.device)
        _lib.check(lib.glnn_sage_forward(
            g.indptr.data_ptr(), int(g.indptr.dtype == torch.int64), g.indices.data_ptr(), n,
            feats.data_ptr(), feats.stride(0), layers, self.num_layers, out.data_ptr(),
            out.stride(0), _fwd_flags(log_softmax, exact), ws.data_ptr(), ws.numel(), _lib.stream()),
            "glnn_sage_forward")
        del keep
        return out


# ------------------------------------------------------------------------------------------------
# GCN (teacher)
# ------------------------------------------------------------------------------------------------
class GraphConv(nn.Module):
    """Parameter holder with dgl.nn.GraphConv's state_dict names and init (rule R3): weight is
    [in, out] (xavier_uniform), bias zeros."""

    def __init__(self, in_feats, out_feats, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.zeros(out_feats))
        nn.init.xavier_uniform_(self.weight)
        self._activation = activation


class GCN(_Encoder):
    def __init__(self, num_layers, input_dim, hidden_dim, output_dim, dropout_ratio, activation,
                 norm_type="none"):
        super().__init__()
        self.dropout = nn.Dropout(dropout_ratio)
        # the reference gives the ONLY layer of a 1-layer GCN the activation too (models.py:168-169):
        # its logits pass through ReLU before log_softmax
        self._build(lambda i, a, b: GraphConv(a, b, activation if (i != num_layers - 1 or num_layers == 1)
                                              else None),
                    num_layers, input_dim, hidden_dim, output_dim, norm_type)

    def forward(self, g, feats, log_softmax=False, exact=None):
        if self.training:
            # train mode (batch statistics, dropout) on the kernels; plain tensors, no autograd --
            # training goes through train_and_eval.train (hand-written backward, teacher_train.py)
            from . import teacher_train
            h_list, h, _ = teacher_train.gcn_forward_train(self, _as_graph(g), feats)
            return h_list, (h.log_softmax(1) if log_softmax else h)
        g = _as_graph(g)
        _lib.require_cuda(feats)
        g = g.to(feats.device)
        if g.has_zero_in_degree():
            raise ValueError("There are 0-in-degree nodes in the graph (dgl GraphConv would raise; "
                             "add self-loops)")
        lib = _lib.load()
        feats = feats if feats.stride(-1) == 1 else feats.contiguous()
        n = g.num_nodes()
        layers, keep = self._gnn_layer_table([l.weight for l in self.layers],
                                             [l.bias for l in self.layers])
        for l, conv in enumerate(self.layers):
            layers[l].d_in, layers[l].d_out = conv.weight.shape
        ns, nd = g.gcn_norms()
        nbytes = lib.glnn_gnn_forward_workspace_bytes(n, layers, self.num_layers)
        ws = self._workspace(nbytes, feats.device)
        out = torch.empty(n, self.layers[-1].weight.shape[1], dtype=torch.float32,
                          device=feats.device)
        _lib.check(lib.glnn_gcn_forward(
            g.indptr.data_ptr(), int(g.indptr.dtype == torch.int64), g.indices.data_ptr(), n,
            ns.data_ptr(), nd.data_ptr(), feats.data_ptr(), feats.stride(0), layers,
            self.num_layers, out.data_ptr(), out.stride(0), _fwd_flags(log_softmax, exact), ws.data_ptr(),
            ws.numel(), _lib.stream()), "glnn_gcn_forward")
        del keep
        return [], out


# ------------------------------------------------------------------------------------------------
# wrapper
# ------------------------------------------------------------------------------------------------
class Model(nn.Module):
    """Same dispatch as the reference: the first of MLP / SAGE / GCN / GAT / APPNP that occurs as a
    substring of conf["model_name"] (so "MLP3w4" and "GA1MLP" are MLPs)."""

    def __init__(self, conf):
        super().__init__()
        name = conf["model_name"]
        self.model_name = name
        common = dict(num_layers=conf["num_layers"], input_dim=conf["feat_dim"],
                      hidden_dim=conf["hidden_dim"], output_dim=conf["label_dim"],
                      dropout_ratio=conf["dropout_ratio"])
        if "MLP" in name:
            enc = MLP(norm_type=conf["norm_type"], **common)
        elif "SAGE" in name:
            enc = SAGE(activation=F.relu, norm_type=conf["norm_type"], **common)
        elif "GCN" in name:
            enc = GCN(activation=F.relu, norm_type=conf["norm_type"], **common)
        elif "GAT" in name or "APPNP" in name:
            raise NotImplementedError(
                f"{name}: GAT/APPNP teachers are outside the B200 hot path (SURVEY.md section 2.1 #5)")
        else:
            raise ValueError(f"unknown model_name {name!r}")
        self.encoder = enc.to(conf["device"])

    def forward(self, data, feats):
        if "MLP" in self.model_name:
            return self.encoder(feats)[1]
        return self.encoder(data, feats)[1]

    def forward_fitnet(self, data, feats):
        if "MLP" in self.model_name:
            return self.encoder._forward_autograd(feats)
        return self.encoder(data, feats)

    def inference(self, data, feats):
        if "SAGE" in self.model_name:
            return self.encoder.inference(data, feats)
        return self.forward(data, feats)
