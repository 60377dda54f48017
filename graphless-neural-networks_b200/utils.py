"""Plumbing with the reference's names and behaviour (/root/reference/utils.py): seeding, YAML config
merge, output-dir helpers, logger, split helpers, evaluator, and the two graph utilities that sit
on the aggregation kernel (feature_prop) or next to it (compute_min_cut_loss)."""
import logging
import os
import random
import shutil
from datetime import datetime

import numpy as np
import torch
import yaml

try:  # the reference uses pytz("US/Pacific"); zoneinfo gives the same wall clock without the dep
    from zoneinfo import ZoneInfo
    _TZ = ZoneInfo("US/Pacific")
except Exception:  # pragma: no cover - tzdata missing
    _TZ = None

CPF_data = ["cora", "citeseer", "pubmed", "a-computer", "a-photo"]
OGB_data = ["ogbn-arxiv", "ogbn-products"]
NonHom_data = ["pokec", "penn94"]
BGNN_data = ["house_class", "vk_class"]


def set_seed(seed):
    """utils.py:19-26."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_training_config(config_path, model_name, dataset):
    """global section overlaid by [dataset][model_name]; sets model_name (utils.py:29-41)."""
    with open(config_path, "r") as fh:
        full = yaml.load(fh, Loader=yaml.FullLoader)
    merged = dict(full["global"])
    specific = full[dataset][model_name]
    if specific is not None:
        merged.update(specific)
    merged["model_name"] = model_name
    return merged


def check_writable(path, overwrite=True):
    if not os.path.exists(path):
        os.makedirs(path)
    elif overwrite:
        shutil.rmtree(path)
        os.makedirs(path)


def check_readable(path):
    if not os.path.exists(path):
        raise ValueError(f"No such file or directory! {path}")


def _pacific_time(*_):
    return datetime.now(_TZ).timetuple()


def get_logger(filename, console_log=False, log_level=logging.INFO):
    """File (and optionally console) logger with '%b%d %H-%M-%S: msg' lines in US/Pacific time."""
    logger = logging.getLogger(__name__)
    logger.propagate = False
    logger.setLevel(log_level)
    for h in list(logger.handlers):
        logger.removeHandler(h)
    fmt = logging.Formatter("%(asctime)s: %(message)s", datefmt="%b%d %H-%M-%S")
    fmt.converter = _pacific_time
    handlers = [logging.FileHandler(filename)] + ([logging.StreamHandler()] if console_log else [])
    for h in handlers:
        h.setFormatter(fmt)
        logger.addHandler(h)
    return logger


def idx_split(idx, ratio, seed=0):
    """Random split of idx into ratio / (1 - ratio) parts (utils.py:88-100)."""
    set_seed(seed)
    n = len(idx)
    cut = int(n * ratio)
    shuffle = torch.randperm(n)
    return idx[shuffle[:cut]], idx[shuffle[cut:]]


def graph_split(idx_train, idx_val, idx_test, rate, seed):
    """Inductive split: hide `rate` of the test nodes (utils.py:103-127).  Returns observed-graph
    indices (obs_*) plus idx_obs / idx_test_ind in original numbering."""
    idx_test_ind, idx_test_tran = idx_split(idx_test, rate, seed)
    idx_obs = torch.cat([idx_train, idx_val, idx_test_tran])
    n1, n2 = idx_train.shape[0], idx_val.shape[0]
    obs_all = torch.arange(idx_obs.shape[0])
    return obs_all[:n1], obs_all[n1:n1 + n2], obs_all[n1 + n2:], idx_obs, idx_test_ind


def get_evaluator(dataset):
    """Plain argmax accuracy for every dataset (the OGB variant in the reference is dead code,
    shadowed by the second definition at utils.py:151-156)."""

    def evaluator(out, labels):
        pred = out.argmax(1)
        return pred.eq(labels).float().mean().item()

    evaluator._glnn_argmax_accuracy = True  # lets evaluate() use the fused reduction kernel
    return evaluator


def compute_min_cut_loss(g, out):
    """trace(S^T A S) / trace(S^T D S) with S = exp(out) (utils.py:159-168), computed sparsely
    (the reference densifies A, which is impossible beyond small graphs)."""
    s = out.detach().to(torch.float64).exp().cpu()
    src, dst = g.to("cpu").edges()
    # A[i, j] = #edges i -> j  =>  trace(S^T A S) = sum_e <S[src_e], S[dst_e]>
    num = (s[src] * s[dst]).sum()
    deg = g.to("cpu").in_degrees().to(torch.float64)
    den = (deg.unsqueeze(1) * s * s).sum()
    return (num / den).item()


def feature_prop(feats, g, k):
    """(D^-1/2 A D^-1/2)^k X, hop by hop (utils.py:171-189) on the aggregation kernel: both degree
    scalings are fused into the gather (src_scale / dst_scale)."""
    from . import ops
    assert feats.shape[0] == g.num_nodes()
    if not feats.is_cuda:
        raise RuntimeError("feature_prop runs on the B200 aggregation kernel; move feats to CUDA")
    g = g.to(feats.device)
    norm = g.in_degrees().to(torch.float32).clamp(min=1).pow(-0.5).contiguous()
    h = feats.float().contiguous()
    for _ in range(k):
        h = ops.spmm_csr(g.indptr, g.indices, h, src_scale=norm, dst_scale=norm)
    return h
