"""Step functions and runners with the signatures of /root/reference/train_and_eval.py, hot loops
routed to libglnn_b200.so:

  train_mini_batch   (ref :59-86)   -> mlp_engine.train_pass  (one fused CUDA-graph pass, a single
                                       host read of the loss per PASS instead of per step)
  evaluate_mini_batch (ref :108-136) -> mlp_engine.eval_forward + fused NLL/accuracy reduction
  evaluate           (ref :89-105)   -> SAGE.inference / GCN.forward kernels + fused reduction
  train / train_sage (ref :12-56)    -> teacher_train: forward + hand-written backward + Adam as a
                                       kernel sequence, device-side neighbour sampling

The runners keep the reference's bookkeeping (early stopping on `>=`, in-memory best state, log
line formats, loss_and_score rows) because the experiment scripts parse them.
"""
import copy
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import mlp_engine, ops
from .graph import CSRGraph, FullNeighborLoader
from .models import GCN, MLP, SAGE
from .utils import set_seed

_warned = set()


def _warn_once(key, msg):
    if key not in _warned:
        _warned.add(key)
        warnings.warn(msg, stacklevel=3)


def torch_fallback_allowed():
    """The step functions have NO silent fallback: a model / criterion / optimizer / device the B200
    kernels do not implement raises.  GLNN_ALLOW_TORCH_FALLBACK=1 is the explicit escape hatch that
    lets the generic torch loop run instead (used by the CPU tests that compare the runners' host
    bookkeeping with the reference's own functions); it is never a measured path."""
    import os
    return os.environ.get("GLNN_ALLOW_TORCH_FALLBACK", "0") not in ("", "0")


def _outside_fused_path(key, what):
    from ._lib import GlnnError
    if not torch_fallback_allowed():
        raise GlnnError(f"{what}: outside the fused B200 path (it needs CUDA tensors, a glnn_b200 MLP / "
                        "SAGE / GCN with norm_type 'none' or 'batch', NLLLoss / KLDivLoss(batchmean, "
                        "log_target) and torch.optim.Adam).  There is no silent fallback; set "
                        "GLNN_ALLOW_TORCH_FALLBACK=1 to run the generic torch loop explicitly.")
    _warn_once(key, f"{what}: model/criterion/optimizer outside the fused B200 path; "
                    "running the generic autograd loop")


# ------------------------------------------------------------------------------------------------
# criterion / evaluator recognition
# ------------------------------------------------------------------------------------------------
def _is_plain_nll(criterion):
    return (type(criterion) is nn.NLLLoss and criterion.reduction == "mean"
            and criterion.weight is None and criterion.ignore_index == -100)


def _is_batchmean_kl(criterion):
    return (type(criterion) is nn.KLDivLoss and criterion.reduction == "batchmean"
            and getattr(criterion, "log_target", False))


def _is_argmax_accuracy(evaluator):
    return getattr(evaluator, "_glnn_argmax_accuracy", False)


def _loss_and_score(out, labels, criterion, evaluator, idx_eval):
    """criterion + evaluator of the reference's evaluate functions; one fused reduction kernel
    when they are the stock NLLLoss / argmax accuracy."""
    if out.is_cuda and _is_plain_nll(criterion) and _is_argmax_accuracy(evaluator) \
            and labels.dtype == torch.int64 and labels.dim() == 1:
        idx = None
        if idx_eval is not None:
            idx = torch.as_tensor(idx_eval, dtype=torch.int64, device=out.device)
        cnt = out.shape[0] if idx is None else idx.numel()
        s = ops.nll_acc(out, labels, idx).tolist()
        return s[0] / cnt, s[1] / cnt
    if idx_eval is None:
        return criterion(out, labels).item(), evaluator(out, labels)
    return criterion(out[idx_eval], labels[idx_eval]).item(), evaluator(out[idx_eval], labels[idx_eval])


# ------------------------------------------------------------------------------------------------
# 1. step functions
# ------------------------------------------------------------------------------------------------
def train(model, data, feats, labels, criterion, optimizer, idx_train, lamb=1):
    """Full-batch GNN step (GCN teacher, ref :12-29): forward, log_softmax + NLL, hand-written
    backward and Adam as one kernel sequence (teacher_train.gcn_train_step) -- no autograd."""
    from . import teacher_train
    if not isinstance(model.encoder, GCN):
        raise NotImplementedError("train(): full-batch training is implemented for the GCN teacher "
                                  "(SAGE trains with train_sage, MLP with train_mini_batch)")
    return teacher_train.gcn_train_step(model, data, feats, labels, criterion, optimizer, idx_train,
                                        lamb).item()


def train_sage(model, dataloader, feats, labels, criterion, optimizer, lamb=1):
    """Sampled-block SAGE training (ref :32-56) -- a 'next' row of the scope table."""
    from . import teacher_train
    return teacher_train.train_sage(model, dataloader, feats, labels, criterion, optimizer, lamb)


def _batch_index(n, batch_size):
    """The reference's batching: CPU randperm, floor(n / bs) full batches (tail dropped), or one
    batch of all n rows when n < bs."""
    num_batches = max(1, n // batch_size)
    idx = torch.randperm(n)[: num_batches * batch_size]
    return idx.view(1, -1) if num_batches == 1 else idx.view(num_batches, batch_size)


def train_mini_batch(model, feats, labels, batch_size, criterion, optimizer, lamb=1):
    model.train()
    idx_batch = _batch_index(feats.shape[0], batch_size)
    num_batches = idx_batch.shape[0]
    enc = model.encoder
    kind_ok = (_is_plain_nll(criterion) and labels.dtype == torch.int64) or \
        (_is_batchmean_kl(criterion) and labels.is_floating_point())
    if isinstance(enc, MLP) and feats.is_cuda and enc.fused_supported() and kind_ok \
            and mlp_engine.optimizer_supported(enc, optimizer) \
            and not (enc.norm_type == "batch" and idx_batch.shape[1] < 2):
        loss_sum = mlp_engine.train_pass(enc, optimizer, feats, labels, idx_batch, lamb)
        return loss_sum.item() / num_batches
    _outside_fused_path("tmb", "train_mini_batch")
    total_loss = 0.0
    for i in range(num_batches):
        rows = idx_batch[i]
        out = model(None, feats[rows]).log_softmax(dim=1)
        loss = criterion(out, labels[rows])
        total_loss += loss.item()
        optimizer.zero_grad()
        (loss * lamb).backward()
        optimizer.step()
    return total_loss / num_batches


def evaluate(model, data, feats, labels, criterion, evaluator, idx_eval=None):
    """-> (log-probabilities of ALL nodes, loss, score); loss/score on idx_eval if given."""
    model.eval()
    with torch.no_grad():
        enc = model.encoder
        if isinstance(enc, SAGE) and feats.is_cuda:
            out = enc.inference(data, feats, log_softmax=True)
        elif isinstance(enc, GCN) and feats.is_cuda:
            out = enc(data, feats, log_softmax=True)[1]
        elif isinstance(enc, MLP) and feats.is_cuda and enc.fused_supported():
            out = mlp_engine.eval_forward(enc, feats, log_softmax=True)
        else:
            _outside_fused_path("ev", "evaluate")
            out = model.inference(data, feats).log_softmax(dim=1)
        loss, score = _loss_and_score(out, labels, criterion, evaluator, idx_eval)
    return out, loss, score


def evaluate_mini_batch(model, feats, labels, criterion, batch_size, evaluator, idx_eval=None):
    model.eval()
    with torch.no_grad():
        enc = model.encoder
        if isinstance(enc, MLP) and feats.is_cuda and enc.fused_supported():
            # eval-mode rows are independent of the batching, so the chunk size only bounds scratch
            out_all = mlp_engine.eval_forward(enc, feats, log_softmax=True)
        else:
            _outside_fused_path("emb", "evaluate_mini_batch")
            chunks = [model.inference(None, feats[s:s + batch_size]).log_softmax(dim=1)
                      for s in range(0, len(feats), batch_size)]
            out_all = torch.cat(chunks)
        loss, score = _loss_and_score(out_all, labels, criterion, evaluator, idx_eval)
    return out_all, loss, score


# ------------------------------------------------------------------------------------------------
# 2./3. runners
# ------------------------------------------------------------------------------------------------
class _BestTracker:
    """Early stopping exactly as the reference does it: improve on `score_val >= best`, keep a deep
    copy of the state_dict in memory, stop after `patience` non-improving evaluations."""

    def __init__(self, model, patience):
        self.model, self.patience = model, patience
        self.best_epoch, self.best_score, self.count, self.state = 0, 0, 0, None

    def update(self, epoch, score_val):
        if score_val >= self.best_score:
            self.best_epoch, self.best_score, self.count = epoch, score_val, 0
            self.state = copy.deepcopy(self.model.state_dict())
        else:
            self.count += 1

    def should_stop(self, epoch, max_epoch):
        return self.count == self.patience or epoch == max_epoch

    def restore(self):
        self.model.load_state_dict(self.state)


def _eval_loader(g, batch_size):
    return FullNeighborLoader(g, batch_size)


def run_transductive(conf, model, g, feats, labels, indices, criterion, evaluator, optimizer, logger,
                     loss_and_score):
    set_seed(conf["seed"])
    device, batch_size = conf["device"], conf["batch_size"]
    idx_train, idx_val, idx_test = indices
    feats, labels = feats.to(device), labels.to(device)
    is_sage, is_mlp = "SAGE" in model.model_name, "MLP" in model.model_name

    if is_sage:
        from . import teacher_train
        g.create_formats_()
        g = g.to(device)
        data = teacher_train.NeighborLoader(g, idx_train, conf["fan_out"], batch_size, shuffle=True)
        data_eval = _eval_loader(g, batch_size)
    elif is_mlp:
        parts = [(feats[i], labels[i]) for i in (idx_train, idx_val, idx_test)]
    else:
        g = g.to(device)
        data = data_eval = g

    best = _BestTracker(model, conf["patience"])
    for epoch in range(1, conf["max_epoch"] + 1):
        if is_sage:
            loss = train_sage(model, data, feats, labels, criterion, optimizer)
        elif is_mlp:
            loss = train_mini_batch(model, parts[0][0], parts[0][1], batch_size, criterion, optimizer)
        else:
            loss = train(model, data, feats, labels, criterion, optimizer, idx_train)

        if epoch % conf["eval_interval"] == 0:
            if is_mlp:
                res = [evaluate_mini_batch(model, f, y, criterion, batch_size, evaluator)[1:]
                       for f, y in parts]
                (loss_train, score_train), (loss_val, score_val), (loss_test, score_test) = res
            else:
                out, loss_train, score_train = evaluate(model, data_eval, feats, labels, criterion,
                                                        evaluator, idx_train)
                loss_val, score_val = _loss_and_score(out, labels, criterion, evaluator, idx_val)
                loss_test, score_test = _loss_and_score(out, labels, criterion, evaluator, idx_test)
            logger.debug(f"Ep {epoch:3d} | loss: {loss:.4f} | s_train: {score_train:.4f} | "
                         f"s_val: {score_val:.4f} | s_test: {score_test:.4f}")
            loss_and_score += [[epoch, loss_train, loss_val, loss_test, score_train, score_val,
                                score_test]]
            best.update(epoch, score_val)
        if best.should_stop(epoch, conf["max_epoch"]):
            break

    best.restore()
    if is_mlp:
        out, _, score_val = evaluate_mini_batch(model, feats, labels, criterion, batch_size, evaluator,
                                                idx_val)
    else:
        out, _, score_val = evaluate(model, data_eval, feats, labels, criterion, evaluator, idx_val)
    score_test = _loss_and_score(out, labels, criterion, evaluator, idx_test)[1]
    logger.info(f"Best valid model at epoch: {best.best_epoch: 3d}, score_val: {score_val :.4f}, "
                f"score_test: {score_test :.4f}")
    return out, score_val, score_test


def run_inductive(conf, model, g, feats, labels, indices, criterion, evaluator, optimizer, logger,
                  loss_and_score):
    """Observed-subgraph training, evaluation on the observed (transductive) and on the hidden
    (inductive) test nodes over the full graph (ref :290-512)."""
    set_seed(conf["seed"])
    device, batch_size = conf["device"], conf["batch_size"]
    obs_idx_train, obs_idx_val, obs_idx_test, idx_obs, idx_test_ind = indices
    feats, labels = feats.to(device), labels.to(device)
    obs_feats, obs_labels = feats[idx_obs], labels[idx_obs]
    is_sage, is_mlp = "SAGE" in model.model_name, "MLP" in model.model_name

    if is_sage or not is_mlp:
        g = g.to(device)
        obs_g = g.subgraph(idx_obs.to(device))
    if is_sage:
        from . import teacher_train
        obs_data = teacher_train.NeighborLoader(obs_g, obs_idx_train, conf["fan_out"], batch_size,
                                                shuffle=True)
        obs_data_eval, data_eval = _eval_loader(obs_g, batch_size), _eval_loader(g, batch_size)
    elif is_mlp:
        parts = [(obs_feats[i], obs_labels[i]) for i in (obs_idx_train, obs_idx_val, obs_idx_test)]
        parts.append((feats[idx_test_ind], labels[idx_test_ind]))
    else:
        obs_data = obs_data_eval = obs_g
        data_eval = g

    best = _BestTracker(model, conf["patience"])
    for epoch in range(1, conf["max_epoch"] + 1):
        if is_sage:
            loss = train_sage(model, obs_data, obs_feats, obs_labels, criterion, optimizer)
        elif is_mlp:
            loss = train_mini_batch(model, parts[0][0], parts[0][1], batch_size, criterion, optimizer)
        else:
            loss = train(model, obs_data, obs_feats, obs_labels, criterion, optimizer, obs_idx_train)

        if epoch % conf["eval_interval"] == 0:
            if is_mlp:
                res = [evaluate_mini_batch(model, f, y, criterion, batch_size, evaluator)[1:]
                       for f, y in parts]
                ((loss_train, score_train), (loss_val, score_val), (loss_test_tran, score_test_tran),
                 (loss_test_ind, score_test_ind)) = res
            else:
                obs_out, loss_train, score_train = evaluate(model, obs_data_eval, obs_feats, obs_labels,
                                                            criterion, evaluator, obs_idx_train)
                loss_val, score_val = _loss_and_score(obs_out, obs_labels, criterion, evaluator,
                                                      obs_idx_val)
                loss_test_tran, score_test_tran = _loss_and_score(obs_out, obs_labels, criterion,
                                                                  evaluator, obs_idx_test)
                # hidden test nodes are evaluated on the full graph
                out, loss_test_ind, score_test_ind = evaluate(model, data_eval, feats, labels,
                                                              criterion, evaluator, idx_test_ind)
            logger.debug(f"Ep {epoch:3d} | loss: {loss:.4f} | s_train: {score_train:.4f} | s_val: "
                         f"{score_val:.4f} | s_tt: {score_test_tran:.4f} | s_ti: {score_test_ind:.4f}")
            loss_and_score += [[epoch, loss_train, loss_val, loss_test_tran, loss_test_ind,
                                score_train, score_val, score_test_tran, score_test_ind]]
            best.update(epoch, score_val)
        if best.should_stop(epoch, conf["max_epoch"]):
            break

    best.restore()
    if is_mlp:
        obs_out, _, score_val = evaluate_mini_batch(model, obs_feats, obs_labels, criterion, batch_size,
                                                    evaluator, obs_idx_val)
        out, _, score_test_ind = evaluate_mini_batch(model, feats, labels, criterion, batch_size,
                                                     evaluator, idx_test_ind)
    else:
        obs_out, _, score_val = evaluate(model, obs_data_eval, obs_feats, obs_labels, criterion,
                                         evaluator, obs_idx_val)
        out, _, score_test_ind = evaluate(model, data_eval, feats, labels, criterion, evaluator,
                                          idx_test_ind)
    score_test_tran = _loss_and_score(obs_out, obs_labels, criterion, evaluator, obs_idx_test)[1]
    out[idx_obs] = obs_out
    logger.info(f"Best valid model at epoch: {best.best_epoch :3d}, score_val: {score_val :.4f}, "
                f"score_test_tran: {score_test_tran :.4f}, score_test_ind: {score_test_ind :.4f}")
    return out, score_val, score_test_tran, score_test_ind


def distill_run_transductive(conf, model, feats, labels, out_t_all, distill_indices, criterion_l,
                             criterion_t, evaluator, optimizer, logger, loss_and_score):
    """Student distillation (ref :520-606): every epoch is a hard-label pass weighted lamb followed
    by a soft-label pass weighted 1-lamb -- two separate sequences of optimizer steps, never one
    summed loss -- then three evaluations."""
    set_seed(conf["seed"])
    device, batch_size, lamb = conf["device"], conf["batch_size"], conf["lamb"]
    idx_l, idx_t, idx_val, idx_test = distill_indices
    feats, labels, out_t_all = feats.to(device), labels.to(device), out_t_all.to(device)
    feats_l, labels_l = feats[idx_l], labels[idx_l]
    feats_t, out_t = feats[idx_t], out_t_all[idx_t]
    feats_val, labels_val = feats[idx_val], labels[idx_val]
    feats_test, labels_test = feats[idx_test], labels[idx_test]

    best = _BestTracker(model, conf["patience"])
    for epoch in range(1, conf["max_epoch"] + 1):
        loss_l = train_mini_batch(model, feats_l, labels_l, batch_size, criterion_l, optimizer, lamb)
        loss_t = train_mini_batch(model, feats_t, out_t, batch_size, criterion_t, optimizer, 1 - lamb)
        loss = loss_l + loss_t
        if epoch % conf["eval_interval"] == 0:
            _, loss_l, score_l = evaluate_mini_batch(model, feats_l, labels_l, criterion_l, batch_size,
                                                     evaluator)
            _, loss_val, score_val = evaluate_mini_batch(model, feats_val, labels_val, criterion_l,
                                                         batch_size, evaluator)
            _, loss_test, score_test = evaluate_mini_batch(model, feats_test, labels_test, criterion_l,
                                                           batch_size, evaluator)
            logger.debug(f"Ep {epoch:3d} | loss: {loss:.4f} | s_l: {score_l:.4f} | s_val: "
                         f"{score_val:.4f} | s_test: {score_test:.4f}")
            loss_and_score += [[epoch, loss_l, loss_val, loss_test, score_l, score_val, score_test]]
            best.update(epoch, score_val)
        if best.should_stop(epoch, conf["max_epoch"]):
            break

    best.restore()
    out, _, score_val = evaluate_mini_batch(model, feats, labels, criterion_l, batch_size, evaluator,
                                            idx_val)
    score_test = _loss_and_score(out, labels, criterion_l, evaluator, idx_test)[1]
    logger.info(f"Best valid model at epoch: {best.best_epoch: 3d}, score_val: {score_val :.4f}, "
                f"score_test: {score_test :.4f}")
    return out, score_val, score_test


def distill_run_inductive(conf, model, feats, labels, out_t_all, distill_indices, criterion_l,
                          criterion_t, evaluator, optimizer, logger, loss_and_score):
    """Inductive distillation (ref :609-742): train on observed nodes, report observed-test and
    hidden-test scores."""
    set_seed(conf["seed"])
    device, batch_size, lamb = conf["device"], conf["batch_size"], conf["lamb"]
    obs_idx_l, obs_idx_t, obs_idx_val, obs_idx_test, idx_obs, idx_test_ind = distill_indices
    feats, labels, out_t_all = feats.to(device), labels.to(device), out_t_all.to(device)
    obs_feats, obs_labels, obs_out_t = feats[idx_obs], labels[idx_obs], out_t_all[idx_obs]
    feats_l, labels_l = obs_feats[obs_idx_l], obs_labels[obs_idx_l]
    feats_t, out_t = obs_feats[obs_idx_t], obs_out_t[obs_idx_t]
    feats_val, labels_val = obs_feats[obs_idx_val], obs_labels[obs_idx_val]
    feats_tt, labels_tt = obs_feats[obs_idx_test], obs_labels[obs_idx_test]
    feats_ti, labels_ti = feats[idx_test_ind], labels[idx_test_ind]

    best = _BestTracker(model, conf["patience"])
    for epoch in range(1, conf["max_epoch"] + 1):
        loss_l = train_mini_batch(model, feats_l, labels_l, batch_size, criterion_l, optimizer, lamb)
        loss_t = train_mini_batch(model, feats_t, out_t, batch_size, criterion_t, optimizer, 1 - lamb)
        loss = loss_l + loss_t
        if epoch % conf["eval_interval"] == 0:
            ev = [evaluate_mini_batch(model, f, y, criterion_l, batch_size, evaluator)[1:]
                  for f, y in ((feats_l, labels_l), (feats_val, labels_val), (feats_tt, labels_tt),
                               (feats_ti, labels_ti))]
            (loss_l, score_l), (loss_val, score_val), (loss_tt, score_tt), (loss_ti, score_ti) = ev
            logger.debug(f"Ep {epoch:3d} | l: {loss:.4f} | s_l: {score_l:.4f} | s_val: {score_val:.4f} "
                         f"| s_tt: {score_tt:.4f} | s_ti: {score_ti:.4f}")
            loss_and_score += [[epoch, loss_l, loss_val, loss_tt, loss_ti, score_l, score_val,
                                score_tt, score_ti]]
            best.update(epoch, score_val)
        if best.should_stop(epoch, conf["max_epoch"]):
            break

    best.restore()
    obs_out, _, score_val = evaluate_mini_batch(model, obs_feats, obs_labels, criterion_l, batch_size,
                                                evaluator, obs_idx_val)
    out, _, score_test_ind = evaluate_mini_batch(model, feats, labels, criterion_l, batch_size,
                                                 evaluator, idx_test_ind)
    score_test_tran = _loss_and_score(obs_out, obs_labels, criterion_l, evaluator, obs_idx_test)[1]
    out[idx_obs] = obs_out
    logger.info(f"Best valid model at epoch: {best.best_epoch: 3d} score_val: {score_val :.4f}, "
                f"score_test_tran: {score_test_tran :.4f}, score_test_ind: {score_test_ind :.4f}")
    return out, score_val, score_test_tran, score_test_ind
