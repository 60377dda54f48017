"""Host side of the fused student step (glnn_mlp_train_pass / glnn_mlp_eval).

Memory layout for HBM: all trainable parameters of the MLP live in ONE flat fp32 buffer, and so do
the gradients and both Adam moments, so the optimizer is a single streaming kernel and the whole
step works on fixed addresses (a requirement for replaying it as a CUDA graph).  The module's
nn.Parameters, BatchNorm buffers and the torch.optim.Adam state tensors are re-pointed to views of
those buffers, which keeps `state_dict()`, `load_state_dict()`, `copy.deepcopy(state_dict())`,
`torch.save` and even a later plain `optimizer.step()` coherent with what the kernels wrote.
"""
import ctypes

import torch

from . import _lib


class _Flat:
    __slots__ = ("desc", "params", "grads", "m", "v", "bn", "nbt", "views", "ws", "ws_rows",
                 "loss", "pass_count", "opt_id", "total", "dp")


class _Dp:
    """Data-parallel state of one MLP: the symmetric region (control area + flat params + flat
    grads) shared with the other ranks of `group` over peer-mapped memory."""
    __slots__ = ("group", "world", "rank", "region", "handle", "ctrl", "flat", "grp", "bs")


def _desc(mlp):
    return _lib.MlpDesc(num_layers=mlp.num_layers, feat_dim=mlp.input_dim,
                        hidden_dim=mlp.hidden_dim if mlp.num_layers > 1 else 1,
                        label_dim=mlp.output_dim, norm=1 if (mlp.norm_type == "batch" and
                                                             mlp.num_layers > 1) else 0,
                        dropout=float(mlp.dropout_ratio), bn_eps=1e-5, bn_momentum=0.1)


def _param_order(mlp):
    """(tensor, numel) in the flat layout of include/glnn_b200.h."""
    seq = []
    for lin in mlp.layers:
        seq += [lin.weight, lin.bias]
    if mlp.norm_type == "batch":
        for bn in mlp.norms:
            seq += [bn.weight, bn.bias]
    return seq


def ensure_flat(mlp):
    """Returns the flat state, (re)building it when parameters were moved (model.to(), first use)."""
    if not mlp.fused_supported():
        raise NotImplementedError("fused MLP path supports norm_type 'none' and 'batch'")
    fl = mlp._flat
    params = _param_order(mlp)
    dev = params[0].device
    if dev.type != "cuda":
        raise _lib.GlnnError("glnn_b200 student kernels need the model on a CUDA device")
    ok = fl is not None and fl.params.device == dev
    if ok:
        for p, (off, n) in zip(params, fl.views):
            if p.data_ptr() != fl.params.data_ptr() + 4 * off or not p.is_contiguous():
                ok = False
                break
    if ok:
        return fl
    lib = _lib.load()
    fl = _Flat()
    fl.desc = _desc(mlp)
    total = lib.glnn_mlp_param_count(ctypes.byref(fl.desc))
    if total != sum(p.numel() for p in params):
        raise _lib.GlnnError("flat layout mismatch between the library and the module")
    fl.total, fl.dp = total, None
    fl.params = torch.empty(total, dtype=torch.float32, device=dev)
    fl.grads = torch.zeros(total, dtype=torch.float32, device=dev)
    fl.m = torch.zeros(total, dtype=torch.float32, device=dev)
    fl.v = torch.zeros(total, dtype=torch.float32, device=dev)
    fl.views = []
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            view = fl.params[off:off + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            fl.views.append((off, n))
            off += n
        nbn = lib.glnn_mlp_bn_stat_count(ctypes.byref(fl.desc))
        fl.bn = torch.empty(max(nbn, 1), dtype=torch.float32, device=dev)
        n_norm = len(mlp.norms) if fl.desc.norm else 0
        fl.nbt = torch.zeros(max(n_norm, 1), dtype=torch.int64, device=dev)
        h = mlp.hidden_dim
        for l in range(n_norm):
            bn = mlp.norms[l]
            for j, name in enumerate(("running_mean", "running_var")):
                view = fl.bn[(2 * l + j) * h:(2 * l + j + 1) * h]
                view.copy_(getattr(bn, name))
                getattr(bn, name).data = view
            view = fl.nbt[l]
            view.copy_(bn.num_batches_tracked)
            bn.num_batches_tracked.data = view
    fl.ws, fl.ws_rows = None, 0
    fl.loss = torch.zeros(1, dtype=torch.float32, device=dev)
    fl.pass_count = 0
    fl.opt_id = None
    mlp._flat = fl
    return fl


def enable_data_parallel(mlp, group=None):
    """Marks the MLP for data-parallel training over `group` (default: the world group).  The global
    batches of train_pass are then split by rows over the ranks; see glnn_mlp_train_pass_dp.  Every
    rank must hold the same parameters and call train_pass with the same arguments."""
    import torch.distributed as dist
    group = group or dist.group.WORLD
    world = dist.get_world_size(group)
    if world > 8:
        raise ValueError("data-parallel student: at most 8 ranks (one NVLink box)")
    mlp._dp_group = group if world > 1 else None
    return mlp


def _ensure_dp(mlp, fl, bs):
    """(Re)homes the flat parameter / gradient buffers in a symmetric region once the global batch
    size is known (the control area holds the BatchNorm exchange buffers, sized by it)."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm
    group = mlp._dp_group
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lib = _lib.load()
    if bs % world:
        raise ValueError(f"data-parallel student: batch size {bs} does not split over {world} ranks")
    ctrl = lib.glnn_mlp_dp_control_bytes(ctypes.byref(fl.desc), bs, world)
    if ctrl < 0:
        raise _lib.GlnnError("glnn_mlp_dp_control_bytes failed")
    dp = fl.dp
    if dp is not None and dp.group is group and dp.ctrl >= ctrl:
        dp.bs = bs
        return dp
    flat = lib.glnn_mlp_dp_flat_count(ctypes.byref(fl.desc), world)
    dev = fl.params.device
    nbytes = ctrl + 2 * 4 * flat
    region = symm.empty(nbytes, dtype=torch.uint8, device=dev)
    handle = symm.rendezvous(region, group)
    _lib.check(lib.glnn_mlp_dp_init(region.data_ptr(), ctrl, _lib.stream()), "glnn_mlp_dp_init")
    new_params = region[ctrl:ctrl + 4 * flat].view(torch.float32)
    new_grads = region[ctrl + 4 * flat:ctrl + 8 * flat].view(torch.float32)
    old_m, old_v = fl.m, fl.v
    with torch.no_grad():
        new_params.zero_()
        new_grads.zero_()
        new_params[:fl.total].copy_(fl.params[:fl.total])
        dist.broadcast(new_params, dist.get_global_rank(group, 0), group=group)  # one model
        fl.params, fl.grads = new_params, new_grads
        fl.m = torch.zeros(flat, dtype=torch.float32, device=dev)
        fl.v = torch.zeros(flat, dtype=torch.float32, device=dev)
        fl.m[:fl.total].copy_(old_m[:fl.total])
        fl.v[:fl.total].copy_(old_v[:fl.total])
        for p, (off, n) in zip(_param_order(mlp), fl.views):
            p.data = fl.params[off:off + n].view(p.shape)
    fl.opt_id = None  # optimizer state views are re-bound to the new moment buffers
    dp = _Dp()
    dp.group, dp.world, dp.rank, dp.region, dp.handle = group, world, rank, region, handle
    dp.ctrl, dp.flat, dp.bs = ctrl, flat, bs
    dp.grp = _lib.DpGroup()
    dp.grp.world, dp.grp.rank, dp.grp.bytes = world, rank, nbytes
    for r, ptr in enumerate(handle.buffer_ptrs):
        dp.grp.base[r] = ptr
    fl.dp = dp
    torch.cuda.synchronize()
    dist.barrier(group=group)  # every rank's control area is initialised before anyone signals
    return dp


def _workspace(fl, rows):
    if fl.ws is None or fl.ws_rows < rows:
        lib = _lib.load()
        nbytes = lib.glnn_mlp_workspace_bytes(ctypes.byref(fl.desc), rows)
        if nbytes < 0:
            raise _lib.GlnnError("glnn_mlp_workspace_bytes failed")
        fl.ws = torch.empty(nbytes, dtype=torch.uint8, device=fl.params.device)
        fl.ws_rows = rows
    return fl.ws


def optimizer_supported(mlp, optimizer):
    if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
        return False
    g = optimizer.param_groups[0]
    if g.get("amsgrad") or g.get("maximize"):
        return False
    mine = _param_order(mlp)
    return len(g["params"]) == len(mine) and {id(p) for p in g["params"]} == {id(p) for p in mine}


def _bind_optimizer(mlp, fl, optimizer):
    """Makes optimizer.state[p] views of the flat moments; returns the common step count."""
    step0 = None
    for p, (off, n) in zip(_param_order(mlp), fl.views):
        st = optimizer.state[p]
        m_view = fl.m[off:off + n].view(p.shape)
        v_view = fl.v[off:off + n].view(p.shape)
        if "exp_avg" in st and st["exp_avg"].data_ptr() != m_view.data_ptr():
            m_view.copy_(st["exp_avg"])
            v_view.copy_(st["exp_avg_sq"])
        elif "exp_avg" not in st and fl.opt_id != id(optimizer):
            m_view.zero_()
            v_view.zero_()
        st["exp_avg"], st["exp_avg_sq"] = m_view, v_view
        if "step" not in st:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
        s = int(st["step"])
        if step0 is None:
            step0 = s
        elif s != step0:
            raise _lib.GlnnError("fused Adam needs one common step count for all parameters")
    fl.opt_id = id(optimizer)
    return step0


def train_pass(mlp, optimizer, feats, targets, idx_batch, lamb, drop_masks=None):
    """One pass of train_mini_batch over batches idx_batch [nb, bs] (int64 rows of feats).
    targets: int64 labels [n] (NLL) or fp32 teacher log-probabilities [n, C] (KL).
    Returns the device scalar holding the SUM over steps of the unscaled mean losses."""
    lib = _lib.load()
    fl = ensure_flat(mlp)
    _lib.require_cuda(feats, targets)
    if feats.dtype != torch.float32 or feats.stride(-1) != 1:
        feats = feats.float().contiguous()
    if targets.dtype == torch.int64:
        kind = 0
        targets = targets.contiguous()
    else:
        kind = 1
        targets = targets.float().contiguous()
        if targets.shape[1] != mlp.output_dim:
            raise ValueError("teacher log-probabilities must be [n, label_dim]")
    nb, bs = idx_batch.shape
    idx_batch = idx_batch.contiguous()
    if idx_batch.dtype != torch.int64:
        idx_batch = idx_batch.long()
    dp = None
    if getattr(mlp, "_dp_group", None) is not None:
        import torch.distributed as dist
        dp = _ensure_dp(mlp, fl, bs)
        idx_batch = idx_batch.to(fl.params.device)
        dist.broadcast(idx_batch, dist.get_global_rank(dp.group, 0), group=dp.group)  # same batches
    step0 = _bind_optimizer(mlp, fl, optimizer)
    g = optimizer.param_groups[0]
    hp = _lib.AdamHParams(lr=float(g["lr"]), beta1=float(g["betas"][0]), beta2=float(g["betas"][1]),
                          eps=float(g["eps"]), weight_decay=float(g["weight_decay"]))
    ws = _workspace(fl, bs if dp is None else bs // dp.world)
    fl.loss.zero_()
    seed = (torch.cuda.initial_seed() * 1000003 + fl.pass_count) & ((1 << 63) - 1)
    fl.pass_count += 1
    if drop_masks is not None:
        _lib.require_cuda(drop_masks)
        if drop_masks.dtype != torch.uint8 or drop_masks.numel() != nb * (mlp.num_layers - 1) * bs * \
                mlp.hidden_dim:
            raise ValueError("drop_masks must be uint8 [nb, L-1, bs, hidden]")
        drop_masks = drop_masks.contiguous()
    max_rows = 1 << 22
    per_call = max(1, max_rows // bs)
    done = 0
    while done < nb:
        cnt = min(per_call, nb - done)
        sub = idx_batch[done:done + cnt]
        msk = None
        if drop_masks is not None:
            msk = drop_masks.view(nb, -1)[done:done + cnt]
        args = (ctypes.byref(fl.desc), fl.params.data_ptr(), fl.grads.data_ptr(), fl.m.data_ptr(),
                fl.v.data_ptr(), fl.bn.data_ptr(), fl.nbt.data_ptr() if fl.desc.norm else None,
                step0 + done, ctypes.byref(hp), feats.data_ptr(), feats.stride(0), targets.data_ptr(),
                kind, sub.data_ptr(), cnt, bs, _lib.ptr(msk), seed, float(lamb), fl.loss.data_ptr(),
                ws.data_ptr(), ws.numel(), _lib.stream())
        if dp is None:
            _lib.check(lib.glnn_mlp_train_pass(*args), "glnn_mlp_train_pass")
        else:
            _lib.check(lib.glnn_mlp_train_pass_dp(ctypes.byref(dp.grp), *args), "glnn_mlp_train_pass_dp")
        done += cnt
    for p in _param_order(mlp):
        optimizer.state[p]["step"] += nb
    if dp is not None:
        import torch.distributed as dist
        # each rank advanced only its slice of the Adam moments and holds its share of the loss
        sl = dp.flat // dp.world
        dist.all_gather_into_tensor(fl.m, fl.m[dp.rank * sl:(dp.rank + 1) * sl], group=dp.group)
        dist.all_gather_into_tensor(fl.v, fl.v[dp.rank * sl:(dp.rank + 1) * sl], group=dp.group)
        dist.all_reduce(fl.loss, group=dp.group)
    return fl.loss


def eval_forward(mlp, feats, rows_per_chunk=None, log_softmax=True):
    """Eval-mode forward of all rows of feats -> [n, C] log-probabilities (or logits)."""
    lib = _lib.load()
    fl = ensure_flat(mlp)
    _lib.require_cuda(feats)
    if feats.dtype != torch.float32 or feats.stride(-1) != 1:
        feats = feats.float().contiguous()
    n = feats.shape[0]
    group = getattr(mlp, "_dp_group", None)
    if group is not None and n >= 4096:
        # evaluate_mini_batch sharded by contiguous node ranges: eval-mode
        # rows are independent, every rank computes its range and the log-probabilities are
        # all-gathered in place
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        rpr = (n + world - 1) // world
        out = torch.empty(world * rpr, mlp.output_dim, dtype=torch.float32, device=feats.device)
        lo, hi = min(n, rank * rpr), min(n, (rank + 1) * rpr)
        if hi > lo:
            rows = min(hi - lo, rows_per_chunk or 65536)
            ws = _workspace(fl, rows)
            mine = out[lo:hi]
            _lib.check(lib.glnn_mlp_eval(ctypes.byref(fl.desc), fl.params.data_ptr(),
                                         fl.bn.data_ptr(), feats[lo:hi].data_ptr(), feats.stride(0),
                                         hi - lo, mine.data_ptr(), mine.stride(0), int(log_softmax),
                                         rows, ws.data_ptr(), ws.numel(), _lib.stream()),
                       "glnn_mlp_eval")
        dist.all_gather_into_tensor(out, out[rank * rpr:(rank + 1) * rpr], group=group)
        return out[:n]
    rows = min(max(n, 1), rows_per_chunk or 65536)
    ws = _workspace(fl, rows)
    out = torch.empty(n, mlp.output_dim, dtype=torch.float32, device=feats.device)
    _lib.check(lib.glnn_mlp_eval(ctypes.byref(fl.desc), fl.params.data_ptr(), fl.bn.data_ptr(),
                                 feats.data_ptr(), feats.stride(0) if n > 1 else feats.shape[1], n,
                                 out.data_ptr(), out.stride(0), int(log_softmax), rows,
                                 ws.data_ptr(), ws.numel(), _lib.stream()), "glnn_mlp_eval")
    return out


def eval_logits(mlp, feats):
    return eval_forward(mlp, feats, log_softmax=False)


def flat_grads(mlp):
    """Gradients of the LAST fused step as {state_dict-style name: tensor view} (diagnostics and
    parity tests; the fused path never materialises p.grad)."""
    fl = ensure_flat(mlp)
    names = []
    for i in range(len(mlp.layers)):
        names += [f"layers.{i}.weight", f"layers.{i}.bias"]
    if mlp.norm_type == "batch":
        for i in range(len(mlp.norms)):
            names += [f"norms.{i}.weight", f"norms.{i}.bias"]
    out = {}
    for name, p, (off, n) in zip(names, _param_order(mlp), fl.views):
        out[name] = fl.grads[off:off + n].view(p.shape)
    return out
