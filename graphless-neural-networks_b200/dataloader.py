"""Dataset loading with the reference's `load_data` contract (/root/reference/dataloader.py:42-58):
returns (g, labels, idx_train, idx_val, idx_test) with node features in g.ndata["feat"], where g is
a glnn_b200 CSRGraph instead of a DGLGraph.  One-time host-side preprocessing, not a hot path; the
graph-structure conventions the kernels depend on are kept exactly:
  * CPF datasets (cora, citeseer, pubmed, a-computer, a-photo): unweighted, symmetrised, self-loop
    free largest connected component, THEN one self-loop per node (normalize_adj adds I), edge
    weights dropped (dataloader.py:82-111);
  * ogbn-arxiv: reverse edges appended WITHOUT de-duplication, self-loops removed, one self-loop
    per node added; ogbn-products untouched (dataloader.py:61-79).
"""
import os
from pathlib import Path

import numpy as np
import scipy.sparse as sp
import torch

from .graph import graph as make_graph
from .utils import BGNN_data, CPF_data, NonHom_data, OGB_data


def load_data(dataset, dataset_path, **kwargs):
    if dataset in CPF_data:
        return load_cpf_data(dataset, dataset_path, kwargs["seed"], kwargs["labelrate_train"],
                             kwargs["labelrate_val"])
    if dataset in OGB_data:
        return load_ogb_data(dataset, dataset_path)
    if dataset in NonHom_data + BGNN_data:
        raise NotImplementedError(
            f"{dataset}: the non-homophilous / BGNN loaders (category_encoders, .mat files) are "
            "outside the B200 hot-path scope; only their graph goes through the kernels")
    raise ValueError(f"Unknown dataset: {dataset}")


def load_out_t(out_t_dir):
    """Teacher log-probabilities written by train_teacher.py: out.npz, key arr_0, float32 [N, C]."""
    return torch.from_numpy(np.load(Path(out_t_dir).joinpath("out.npz"))["arr_0"])


# ------------------------------------------------------------------------------------------------
# OGB
# ------------------------------------------------------------------------------------------------
def arxiv_edges(src, dst, n):
    """The reference's ogbn-arxiv preprocessing (dataloader.py:74-77): `g.add_edges(dsts, srcs)`
    appends every edge reversed WITHOUT de-duplication (an edge present in both directions becomes
    two parallel edges each way), `remove_self_loop()` drops every u -> u, `add_self_loop()` adds
    exactly one per node.  Multiplicities are kept: the aggregation counts parallel edges."""
    src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    keep = src != dst
    loops = torch.arange(n, dtype=src.dtype)
    return torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])


def load_ogb_data(dataset, dataset_path):
    try:
        from ogb.nodeproppred import NodePropPredDataset
    except ImportError as e:  # pragma: no cover - ogb is not installed in the build container
        raise ImportError("loading ogbn-* needs the `ogb` package (and the downloaded dataset)") from e
    data = NodePropPredDataset(dataset, dataset_path)
    split = data.get_idx_split()
    graph_dict, labels = data[0]
    src, dst = (torch.from_numpy(graph_dict["edge_index"][i]).long() for i in (0, 1))
    n = int(graph_dict["num_nodes"])
    if dataset == "ogbn-arxiv":
        src, dst = arxiv_edges(src, dst, n)
    g = make_graph((src, dst), num_nodes=n)
    g.ndata["feat"] = torch.from_numpy(graph_dict["node_feat"]).float()
    labels = torch.from_numpy(labels).squeeze().long()
    return g, labels, *(torch.from_numpy(split[k]).long() for k in ("train", "valid", "test"))


# ------------------------------------------------------------------------------------------------
# CPF (.npz in the gnn-benchmark sparse format)
# ------------------------------------------------------------------------------------------------
def _load_npz(path):
    with np.load(path, allow_pickle=True) as z:
        z = dict(z)
    adj = sp.csr_matrix((z["adj_data"], z["adj_indices"], z["adj_indptr"]), shape=z["adj_shape"])
    if "attr_data" in z:
        attr = sp.csr_matrix((z["attr_data"], z["attr_indices"], z["attr_indptr"]),
                             shape=z["attr_shape"])
    elif "attr_matrix" in z:
        attr = z["attr_matrix"]
    else:
        raise ValueError(f"{path}: no node attributes")
    if "labels" not in z:
        raise ValueError(f"{path}: only single-label datasets are supported")
    return adj, attr, np.asarray(z["labels"])


def _standardize(adj, attr, labels):
    """Unweighted + undirected + no self loops, restricted to the largest connected component
    (nodes keep their relative order)."""
    adj = adj.tocsr().astype(np.float32)
    adj.data[:] = 1.0
    adj = adj + adj.T
    adj.data[:] = 1.0
    adj = adj.tolil()
    adj.setdiag(0)
    adj = adj.tocsr()
    adj.eliminate_zeros()
    _, comp = sp.csgraph.connected_components(adj)
    biggest = np.argmax(np.bincount(comp))
    keep = np.flatnonzero(comp == biggest)
    return adj[keep][:, keep], attr[keep], labels[keep]


def _per_class_sample(rng, onehot, per_class, forbidden=None):
    forbidden = set() if forbidden is None else set(int(i) for i in forbidden)
    picks = []
    for c in range(onehot.shape[1]):
        cand = [i for i in np.flatnonzero(onehot[:, c] > 0) if i not in forbidden]
        picks.append(rng.choice(cand, per_class, replace=False))
    return np.concatenate(picks)


def load_cpf_data(dataset, dataset_path, seed, labelrate_train, labelrate_val):
    path = Path.cwd().joinpath(dataset_path, f"{dataset}.npz")
    if not os.path.isfile(path):
        raise ValueError(f"{path} doesn't exist.")
    adj, attr, labels = _standardize(*_load_npz(path))
    classes = np.unique(labels)
    onehot = (labels[:, None] == classes[None, :]).astype(np.float32)
    # same RandomState call sequence as the reference's split sampler: train per class, then val
    # per class avoiding train, test = everything else (dataloader.py:593-702)
    rng = np.random.RandomState(seed)
    idx_train = _per_class_sample(rng, onehot, labelrate_train)
    idx_val = _per_class_sample(rng, onehot, labelrate_val, forbidden=idx_train)
    idx_test = np.setdiff1d(np.arange(labels.shape[0]), np.concatenate([idx_train, idx_val]))
    feats = torch.FloatTensor(np.asarray(attr.todense() if sp.issparse(attr) else attr))
    y = torch.LongTensor(onehot.argmax(axis=1))
    coo = (adj + sp.eye(adj.shape[0], dtype=np.float32)).tocoo()  # normalize_adj adds I; weights dropped
    g = make_graph((coo.row.astype(np.int64), coo.col.astype(np.int64)), num_nodes=adj.shape[0])
    g.ndata["feat"] = feats
    return g, y, torch.LongTensor(idx_train), torch.LongTensor(idx_val), torch.LongTensor(idx_test)
