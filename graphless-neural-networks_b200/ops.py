"""Tensor-level wrappers over the C ABI (one function per entry point of include/glnn_b200.h).
Inputs must be CUDA fp32 tensors with unit stride in the last dimension; the row stride is passed
through as the leading dimension, so column-sliced views work without copies."""
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream


def _ld(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("expected a 2-D tensor with unit stride in the last dimension")
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def _f32(*ts):
    for t in ts:
        if t is not None and t.dtype != torch.float32:
            raise ValueError("glnn_b200 kernels are fp32")


def spmm_csr(indptr, indices, x, d=None, out=None, self_add=False, mean_plus_one=False,
             src_scale=None, dst_scale=None, bias=None, col_scale=None, col_shift=None, relu=0):
    """glnn_spmm_csr_f32.  indptr int32/int64 [n_dst+1], indices int32 [nnz], x [n_src, >=d]."""
    lib = _lib.load()
    require_cuda(indptr, indices, x, out, src_scale, dst_scale, bias, col_scale, col_shift)
    _f32(x, out, src_scale, dst_scale, bias, col_scale, col_shift)
    if indices.dtype != torch.int32:
        raise ValueError("indices must be int32")
    if indptr.dtype not in (torch.int32, torch.int64):
        raise ValueError("indptr must be int32 or int64")
    n_dst = indptr.numel() - 1
    d = x.shape[1] if d is None else d
    if out is None:
        out = torch.empty(n_dst, d, dtype=torch.float32, device=x.device)
    check(lib.glnn_spmm_csr_f32(ptr(indptr), int(indptr.dtype == torch.int64), ptr(indices), ptr(x),
                                _ld(x), ptr(out), _ld(out), n_dst, x.shape[0], d, int(self_add),
                                int(mean_plus_one), ptr(src_scale), ptr(dst_scale), ptr(bias),
                                ptr(col_scale), ptr(col_shift), int(relu), stream()),
          "glnn_spmm_csr_f32")
    return out


def gemm(a, b, trans_a=False, trans_b=False, out=None, row_scale=None, bias=None, col_scale=None,
         col_shift=None, relu=0, impl=0):
    """glnn_gemm_f32: out = epilogue(op(a) @ op(b)).  trans_b=True takes an nn.Linear weight."""
    lib = _lib.load()
    require_cuda(a, b, out, row_scale, bias, col_scale, col_shift)
    _f32(a, b, out, row_scale, bias, col_scale, col_shift)
    m, k = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    kb, n = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if k != kb:
        raise ValueError(f"gemm: inner dimensions differ ({k} vs {kb})")
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    check(lib.glnn_gemm_f32(ptr(a), _ld(a), int(trans_a), ptr(b), _ld(b), int(trans_b), ptr(out),
                            _ld(out), m, n, k, ptr(row_scale), ptr(bias), ptr(col_scale),
                            ptr(col_shift), int(relu), int(impl), stream()), "glnn_gemm_f32")
    return out


def bn_fold(gamma, beta, mean, var, eps):
    lib = _lib.load()
    require_cuda(gamma, beta, mean, var)
    n = gamma.numel()
    out = torch.empty(2, n, dtype=torch.float32, device=gamma.device)
    check(lib.glnn_bn_fold_f32(ptr(gamma), ptr(beta), ptr(mean), ptr(var), float(eps), ptr(out[0]),
                               ptr(out[1]), n, stream()), "glnn_bn_fold_f32")
    return out[0], out[1]


def log_softmax(x, out=None):
    lib = _lib.load()
    require_cuda(x, out)
    _f32(x, out)
    if out is None:
        out = torch.empty(x.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
    check(lib.glnn_log_softmax_f32(ptr(x), _ld(x), ptr(out), _ld(out), x.shape[0], x.shape[1],
                                   stream()), "glnn_log_softmax_f32")
    return out


def nll_acc(logp, labels, idx=None):
    """(sum of -logp[i, y_i], number of argmax hits) over rows idx (None = all) as a 2-float tensor."""
    lib = _lib.load()
    require_cuda(logp, labels, idx)
    if labels.dtype != torch.int64 or (idx is not None and idx.dtype != torch.int64):
        raise ValueError("labels / idx must be int64")
    n = logp.shape[0] if idx is None else idx.numel()
    out = torch.zeros(2, dtype=torch.float32, device=logp.device)
    check(lib.glnn_nll_acc_f32(ptr(logp), _ld(logp), logp.shape[1], ptr(labels), ptr(idx), n,
                               ptr(out), stream()), "glnn_nll_acc_f32")
    return out


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8,
              weight_decay=0.0):
    """glnn_adam_step_f32: in-place torch.optim.Adam update of a flat fp32 buffer (step is 1-based)."""
    lib = _lib.load()
    require_cuda(params, grads, exp_avg, exp_avg_sq)
    _f32(params, grads, exp_avg, exp_avg_sq)
    n = params.numel()
    for t in (params, grads, exp_avg, exp_avg_sq):
        if t.numel() != n or not t.is_contiguous():
            raise ValueError("adam_step: contiguous buffers of equal length expected")
    import ctypes
    hp = _lib.AdamHParams(lr=float(lr), beta1=float(betas[0]), beta2=float(betas[1]), eps=float(eps),
                          weight_decay=float(weight_decay))
    check(lib.glnn_adam_step_f32(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), n, int(step),
                                 ctypes.byref(hp), stream()), "glnn_adam_step_f32")
    return params


def peer_push(src, dsts, ctas=0):
    """glnn_peer_push: contiguous `src` -> every tensor of `dsts` (peer-mapped views of the same
    shape), one SM-driven kernel on the current stream."""
    import ctypes
    lib = _lib.load()
    if not src.is_contiguous() or any(not d.is_contiguous() for d in dsts):
        raise ValueError("peer_push: contiguous slabs expected")
    nbytes = src.numel() * src.element_size()
    arr = (ctypes.c_void_p * max(len(dsts), 1))(*[d.data_ptr() for d in dsts])
    check(lib.glnn_peer_push(ptr(src), arr, len(dsts), nbytes, int(ctas), stream()), "glnn_peer_push")


def nll_loss_grad(logits, labels, rows=None, label_rows=None, lamb=1.0, dlogits=None, loss_out=None):
    """glnn_nll_loss_grad_f32 -> (dlogits, loss_out).  rows / label_rows: int64 selections (see the
    header); dlogits is zero-initialised here when a subset of the rows is selected."""
    lib = _lib.load()
    require_cuda(logits, labels, rows, label_rows, dlogits, loss_out)
    _f32(logits, dlogits, loss_out)
    for t in (labels, rows, label_rows):
        if t is not None and t.dtype != torch.int64:
            raise ValueError("labels / rows / label_rows must be int64")
    m = logits.shape[0] if rows is None else rows.numel()
    if label_rows is not None and label_rows.numel() != m:
        raise ValueError("label_rows must have one entry per selected row")
    if dlogits is None:
        mk = torch.empty if rows is None else torch.zeros
        dlogits = mk(logits.shape[0], logits.shape[1], dtype=torch.float32, device=logits.device)
    if loss_out is None:
        loss_out = torch.zeros(1, dtype=torch.float32, device=logits.device)
    check(lib.glnn_nll_loss_grad_f32(ptr(logits), _ld(logits), logits.shape[1], ptr(labels), ptr(rows),
                                     ptr(label_rows), m, float(lamb), ptr(dlogits), _ld(dlogits),
                                     ptr(loss_out), stream()), "glnn_nll_loss_grad_f32")
    return dlogits, loss_out


class ActBlock:
    """[BatchNorm1d] -> [ReLU] -> [Dropout] in train mode (glnn_act_train_fwd_f32 / _bwd_f32): keeps
    what the backward needs (the input, the saved batch statistics, the dropout seed / mask)."""

    def __init__(self, x, bn=None, relu_post=False, relu_input=False, p_drop=0.0, seed=0, keep_mask=None):
        import ctypes
        require_cuda(x, keep_mask)
        _f32(x)
        self.x, self.bn, self.keep_mask = x, bn, keep_mask
        n, d = x.shape
        q = _lib.ActDesc()
        q.n, q.d, q.relu_post, q.relu_input = n, d, int(relu_post), int(relu_input)
        q.X, q.ldx = ptr(x), _ld(x)
        q.p_drop, q.seed = float(p_drop), int(seed) & ((1 << 64) - 1)
        if keep_mask is not None:
            if keep_mask.dtype != torch.uint8 or keep_mask.numel() != n * d or not keep_mask.is_contiguous():
                raise ValueError("keep_mask must be a contiguous uint8 [n, d] tensor")
            q.keep_mask = ptr(keep_mask)
        self.scratch = None
        if bn is not None:
            self.stats = torch.empty(2, d, dtype=torch.float32, device=x.device)
            self.scratch = torch.empty(2, d, dtype=torch.float32, device=x.device)
            q.gamma, q.beta = ptr(bn.weight), ptr(bn.bias)
            q.running_mean, q.running_var = ptr(bn.running_mean), ptr(bn.running_var)
            q.save_mean, q.save_invstd = ptr(self.stats[0]), ptr(self.stats[1])
            q.eps, q.momentum = float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1)
        self.q = q
        self._byref = ctypes.byref

    def forward(self):
        lib = _lib.load()
        y = torch.empty(self.x.shape[0], self.x.shape[1], dtype=torch.float32, device=self.x.device)
        self.q.Y, self.q.ldy = ptr(y), _ld(y)
        check(lib.glnn_act_train_fwd_f32(self._byref(self.q), ptr(self.scratch), stream()),
              "glnn_act_train_fwd_f32")
        if self.bn is not None:
            self.bn.num_batches_tracked += 1
        return y

    def backward(self, dy, want_dbias=True):
        """-> (dX, dgamma, dbeta, dbias)."""
        lib = _lib.load()
        require_cuda(dy)
        _f32(dy)
        n, d = self.x.shape
        dev = self.x.device
        dx = torch.empty(n, d, dtype=torch.float32, device=dev)
        dg = db = None
        if self.bn is not None:
            dg = torch.empty(d, dtype=torch.float32, device=dev)
            db = torch.empty(d, dtype=torch.float32, device=dev)
        dbias = torch.empty(d, dtype=torch.float32, device=dev) if want_dbias else None
        check(lib.glnn_act_train_bwd_f32(self._byref(self.q), ptr(dy), _ld(dy), ptr(dx), _ld(dx), ptr(dg),
                                         ptr(db), ptr(dbias), ptr(self.scratch), stream()),
              "glnn_act_train_bwd_f32")
        return dx, dg, db, dbias


def spmm_scatter(indptr, indices, dy, scale, dx, self_add=False):
    """glnn_spmm_csr_scatter_f32: dx[u] += scale[v] * dy[v] over the edges u -> v (dx pre-initialised)."""
    lib = _lib.load()
    require_cuda(indptr, indices, dy, scale, dx)
    _f32(dy, scale, dx)
    check(lib.glnn_spmm_csr_scatter_f32(ptr(indptr), int(indptr.dtype == torch.int64), ptr(indices), ptr(dy),
                                        _ld(dy), ptr(scale), ptr(dx), _ld(dx), indptr.numel() - 1,
                                        dy.shape[1], int(self_add), stream()), "glnn_spmm_csr_scatter_f32")
    return dx


def sample_neighbors(indptr, indices, seeds, fanout, rng_seed):
    """glnn_sample_count + glnn_sample_neighbors -> (block indptr int64 [m+1], sampled global source
    ids int32 [total]) for int64 `seeds`."""
    lib = _lib.load()
    require_cuda(indptr, indices, seeds)
    if seeds.dtype != torch.int64:
        raise ValueError("seeds must be int64")
    m = seeds.numel()
    i64 = int(indptr.dtype == torch.int64)
    ptr_out = torch.zeros(m + 1, dtype=torch.int64, device=seeds.device)
    check(lib.glnn_sample_count(ptr(indptr), i64, ptr(seeds), m, int(fanout), ptr(ptr_out[1:]), stream()),
          "glnn_sample_count")
    ptr_out[1:].cumsum_(0)
    total = int(ptr_out[-1])
    src = torch.empty(total, dtype=torch.int32, device=seeds.device)
    check(lib.glnn_sample_neighbors(ptr(indptr), i64, ptr(indices), ptr(seeds), m, int(fanout),
                                    int(rng_seed) & ((1 << 64) - 1), ptr(ptr_out), ptr(src), stream()),
          "glnn_sample_neighbors")
    return ptr_out, src


def block_mark(src, flag):
    lib = _lib.load()
    check(lib.glnn_block_mark(ptr(src), src.numel(), ptr(flag), stream()), "glnn_block_mark")


def block_relabel(src, node_map):
    lib = _lib.load()
    check(lib.glnn_block_relabel(ptr(src), src.numel(), ptr(node_map), stream()), "glnn_block_relabel")
    return src


def csr_from_coo(src, dst, num_nodes, want_out_deg=True):
    """glnn_csr_from_coo: CUDA edge list (int64 or int32) -> (indptr int32 [n+1], indices int32 [E],
    out_deg int64 [n] | None), stable by destination.  Raises ValueError for ids outside [0, n)."""
    lib = _lib.load()
    require_cuda(src, dst)
    if src.dtype != dst.dtype or src.dtype not in (torch.int64, torch.int32):
        raise ValueError("src / dst must both be int64 or both int32")
    if src.dim() != 1 or src.shape != dst.shape:
        raise ValueError("src / dst must be 1-d arrays of the same length")
    src, dst = src.contiguous(), dst.contiguous()
    n, e, dev = int(num_nodes), src.numel(), src.device
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(e, dtype=torch.int32, device=dev)
    out_deg = torch.empty(n, dtype=torch.int64, device=dev) if want_out_deg else None
    status = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.glnn_csr_build_workspace_bytes(n, e))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    check(lib.glnn_csr_from_coo(ptr(src), ptr(dst), int(src.dtype == torch.int64), e, n, ptr(indptr),
                                ptr(indices), ptr(out_deg) if want_out_deg else None, ptr(status),
                                ws.data_ptr() + off, ws_bytes, stream()), "glnn_csr_from_coo")
    bad = int(status.item())
    if bad:
        raise ValueError(f"csr_from_coo: {bad} edges have a node id outside [0, {n})")
    return indptr, indices, out_deg


def csr_subgraph(indptr, indices, relabel, n_new):
    """glnn_csr_subgraph: (new_indptr int32 [n_new+1], new_indices int32, new_out_deg int64 [n_new]) of
    the node-induced subgraph; relabel int32 [n] holds the new id of every kept node, -1 elsewhere."""
    lib = _lib.load()
    require_cuda(indptr, indices, relabel)
    if relabel.dtype != torch.int32 or indices.dtype != torch.int32:
        raise ValueError("relabel / indices must be int32")
    n, n_new, dev = relabel.numel(), int(n_new), relabel.device
    i64 = int(indptr.dtype == torch.int64)
    new_ptr = torch.empty(n_new + 1, dtype=torch.int32, device=dev)
    ws_bytes = 12 * (n_new + 1) + 1024
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    wp = ws.data_ptr() + (-ws.data_ptr()) % 256
    check(lib.glnn_csr_subgraph(ptr(indptr), i64, ptr(indices), n, ptr(relabel), n_new, ptr(new_ptr), None,
                                None, wp, ws_bytes, stream()), "glnn_csr_subgraph")
    total = int(new_ptr[-1].item()) if n_new else 0
    new_idx = torch.empty(total, dtype=torch.int32, device=dev)
    out_deg = torch.empty(n_new, dtype=torch.int64, device=dev)
    check(lib.glnn_csr_subgraph(ptr(indptr), i64, ptr(indices), n, ptr(relabel), n_new, ptr(new_ptr),
                                ptr(new_idx), ptr(out_deg), wp, ws_bytes, stream()), "glnn_csr_subgraph")
    return new_ptr, new_idx, out_deg


class Planes:
    """fp32 matrix kept as bf16 hi / lo planes (see include/glnn_b200.h): .hi/.lo int16 tensors
    [rows, ldp], logical width .cols."""

    def __init__(self, hi, lo, cols):
        self.hi, self.lo, self.cols = hi, lo, cols

    @property
    def shape(self):
        return (self.hi.shape[0], self.cols)

    def float(self):
        f = lambda t: (t.view(torch.int16).to(torch.int32) << 16).view(torch.float32)
        return (f(self.hi) + f(self.lo))[:, :self.cols]


def split_planes(x):
    """glnn_split_planes_f32: fp32 [rows, cols] -> Planes with ldp = cols rounded up to 8."""
    lib = _lib.load()
    require_cuda(x)
    _f32(x)
    rows, cols = x.shape
    ldp = (cols + 7) // 8 * 8
    hi = torch.empty(rows, ldp, dtype=torch.int16, device=x.device)
    lo = torch.empty(rows, ldp, dtype=torch.int16, device=x.device)
    check(lib.glnn_split_planes_f32(ptr(x), _ld(x), rows, cols, ptr(hi), ptr(lo), ldp, stream()),
          "glnn_split_planes_f32")
    return Planes(hi, lo, cols)


def gemm_planes(a, b, trans_a=False, trans_b=False, out=None, out_planes=False, row_scale=None,
                bias=None, col_scale=None, col_shift=None, relu=0):
    """glnn_gemm_bf16x3_planes on Planes operands.  Returns fp32 C, or Planes when out_planes."""
    lib = _lib.load()
    m, k = (a.cols, a.hi.shape[0]) if trans_a else (a.hi.shape[0], a.cols)
    kb, n = (b.cols, b.hi.shape[0]) if trans_b else (b.hi.shape[0], b.cols)
    if k != kb:
        raise ValueError(f"gemm_planes: inner dimensions differ ({k} vs {kb})")
    dev = a.hi.device
    c = ch = cl = None
    ldcp = 0
    if isinstance(out_planes, Planes):  # preallocated (pad columns must already be zero)
        ch, cl, ldcp = out_planes.hi, out_planes.lo, out_planes.hi.stride(0)
    elif out_planes:
        ldcp = (n + 7) // 8 * 8
        ch = torch.zeros(m, ldcp, dtype=torch.int16, device=dev)
        cl = torch.zeros(m, ldcp, dtype=torch.int16, device=dev)
    else:
        c = out if out is not None else torch.empty(m, n, dtype=torch.float32, device=dev)
    check(lib.glnn_gemm_bf16x3_planes(ptr(a.hi), ptr(a.lo), a.hi.stride(0), int(trans_a), ptr(b.hi),
                                      ptr(b.lo), b.hi.stride(0), int(trans_b), ptr(c),
                                      0 if c is None else _ld(c), ptr(ch), ptr(cl), ldcp, m, n, k,
                                      ptr(row_scale), ptr(bias), ptr(col_scale), ptr(col_shift),
                                      int(relu), stream()), "glnn_gemm_bf16x3_planes")
    if isinstance(out_planes, Planes):
        return out_planes
    return Planes(ch, cl, n) if out_planes else c


def spmm_csr_planes(indptr, indices, x, d=None, self_add=False, mean_plus_one=False, src_scale=None,
                    dst_scale=None, bias=None, col_scale=None, col_shift=None, relu=0, out=None):
    """glnn_spmm_csr_planes: the aggregation with its result written as Planes (operand of a
    following tensor-core projection)."""
    lib = _lib.load()
    require_cuda(indptr, indices, x, src_scale, dst_scale, bias, col_scale, col_shift)
    _f32(x, src_scale, dst_scale, bias, col_scale, col_shift)
    n_dst = indptr.numel() - 1
    d = x.shape[1] if d is None else d
    if out is None:
        ldp = (d + 7) // 8 * 8
        out = Planes(torch.zeros(n_dst, ldp, dtype=torch.int16, device=x.device),
                     torch.zeros(n_dst, ldp, dtype=torch.int16, device=x.device), d)
    check(lib.glnn_spmm_csr_planes(ptr(indptr), int(indptr.dtype == torch.int64), ptr(indices), ptr(x),
                                   _ld(x), ptr(out.hi), ptr(out.lo), out.hi.stride(0), n_dst,
                                   x.shape[0], d, int(self_add), int(mean_plus_one), ptr(src_scale),
                                   ptr(dst_scale), ptr(bias), ptr(col_scale), ptr(col_shift),
                                   int(relu), stream()), "glnn_spmm_csr_planes")
    return out


class Q24:
    """24-bit row-packed matrix (see include/glnn_b200.h): .data uint8 [rows, ldq]; a row holds dq
    hi16 values then dq mid8 values (dq = cols rounded up to 8) and pad bytes up to ldq."""

    def __init__(self, data, cols):
        self.data, self.cols = data, cols

    @staticmethod
    def row_bytes(cols):
        return (3 * ((cols + 7) // 8 * 8) + 31) // 32 * 32

    @staticmethod
    def empty(rows, cols, device, zero=False):
        mk = torch.zeros if zero else torch.empty
        return Q24(mk(rows, Q24.row_bytes(cols), dtype=torch.uint8, device=device), cols)

    @property
    def shape(self):
        return (self.data.shape[0], self.cols)

    @property
    def ldq(self):
        return self.data.stride(0)

    def float(self):
        dq = (self.cols + 7) // 8 * 8
        hi = self.data[:, :2 * dq].contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
        mid = self.data[:, 2 * dq:3 * dq].to(torch.int32)
        return ((hi << 16) | (mid << 8)).view(torch.float32)[:, :self.cols]


def quantize_q24(x, out=None):
    """glnn_quantize_q24_f32: fp32 [rows, cols] -> Q24."""
    lib = _lib.load()
    require_cuda(x)
    _f32(x)
    rows, cols = x.shape
    if out is None:
        out = Q24.empty(rows, cols, x.device)
    check(lib.glnn_quantize_q24_f32(ptr(x), _ld(x), rows, cols, ptr(out.data), out.ldq, stream()),
          "glnn_quantize_q24_f32")
    return out


def gemm_planes_q24(a, b, trans_b=True, out=None, row_scale=None, bias=None, col_scale=None,
                    col_shift=None, relu=0):
    """glnn_gemm_bf16x3_planes_q24: projection of Planes operands written as Q24 (N % 8 == 0)."""
    lib = _lib.load()
    m, k = a.hi.shape[0], a.cols
    kb, n = (b.cols, b.hi.shape[0]) if trans_b else (b.hi.shape[0], b.cols)
    if k != kb:
        raise ValueError("gemm_planes_q24: inner dimensions differ")
    if out is None:
        out = Q24.empty(m, n, a.hi.device)
    check(lib.glnn_gemm_bf16x3_planes_q24(ptr(a.hi), ptr(a.lo), a.hi.stride(0), ptr(b.hi), ptr(b.lo),
                                          b.hi.stride(0), int(trans_b), ptr(out.data), out.ldq, m, n, k,
                                          ptr(row_scale), ptr(bias), ptr(col_scale), ptr(col_shift),
                                          int(relu), stream()), "glnn_gemm_bf16x3_planes_q24")
    return out


def spmm(indptr, indices, x, d=None, out=None, out_planes=None, self_add=False, mean_plus_one=False,
         src_scale=None, dst_scale=None, bias=None, col_scale=None, col_shift=None, relu=0,
         log_softmax=0, hot_below=0, s24=None, acc_init=None):
    """glnn_spmm_csr (general form).  x: fp32 tensor [n_src, >= d] or Q24; result in `out` (fp32
    tensor) and/or `out_planes` (Planes); if neither is given an fp32 tensor is allocated.
    log_softmax = c > 0 ends the epilogue with log_softmax over the first c columns (fp32 out
    [n_dst, c]); hot_below = k marks source ids < k as L2-resident (a pure performance hint).
    s24 = S24 copy of the Q24 matrix x (EXPERIMENTAL, compact_s24): glnn_spmm_csr_s24."""
    lib = _lib.load()
    require_cuda(indptr, indices, out, src_scale, dst_scale, bias, col_scale, col_shift)
    _f32(out, src_scale, dst_scale, bias, col_scale, col_shift)
    if indices.dtype != torch.int32:
        raise ValueError("indices must be int32")
    if indptr.dtype not in (torch.int32, torch.int64):
        raise ValueError("indptr must be int32 or int64")
    q = _lib.SpmmDesc()
    q.indptr, q.indices, q.indptr64 = ptr(indptr), ptr(indices), int(indptr.dtype == torch.int64)
    n_dst = indptr.numel() - 1
    q.n_dst = n_dst
    if isinstance(x, Q24):
        require_cuda(x.data)
        q.X_q24, q.ldq, q.n_src = ptr(x.data), x.ldq, x.data.shape[0]
        d = x.cols if d is None else d
        dev, dout = x.data.device, (d + 7) // 8 * 8
    else:
        require_cuda(x)
        _f32(x)
        q.X, q.ldx, q.n_src = ptr(x), _ld(x), x.shape[0]
        d = x.shape[1] if d is None else d
        dev, dout = x.device, d
    q.d = d
    if out is None and out_planes is None:
        out = torch.empty(n_dst, log_softmax if log_softmax else dout, dtype=torch.float32, device=dev)
    if out is not None:
        q.Y, q.ldy = ptr(out), _ld(out)
    if out_planes is not None:
        q.Y_hi, q.Y_lo, q.ldyp = ptr(out_planes.hi), ptr(out_planes.lo), out_planes.hi.stride(0)
    q.self_add, q.mean_plus_one = int(self_add), int(mean_plus_one)
    q.src_scale, q.dst_scale, q.bias = ptr(src_scale), ptr(dst_scale), ptr(bias)
    q.col_scale, q.col_shift, q.relu = ptr(col_scale), ptr(col_shift), int(relu)
    q.log_softmax, q.hot_below = int(log_softmax), int(hot_below)
    if acc_init is not None:   # accumulators start from this fp32 [n_dst, >= d] matrix
        require_cuda(acc_init)
        _f32(acc_init)
        q.Y_init, q.ldyi = ptr(acc_init), _ld(acc_init)
    import ctypes
    if s24 is not None:
        check(lib.glnn_spmm_csr_s24(ctypes.byref(q), ptr(s24.data), s24.data.stride(0), ptr(s24.cap),
                                    stream()), "glnn_spmm_csr_s24")
    else:
        check(lib.glnn_spmm_csr(ctypes.byref(q), stream()), "glnn_spmm_csr")
    return out if out is not None else out_planes


def exp_spmm_tma(indptr, indices, xq, out_planes, stages=8, self_add=True, mean_plus_one=True):
    """glnn_exp_spmm_tma_q24 (EXPERIMENT, tools/exp_spmm_tma.py): 256-wide q24 aggregation with the
    neighbour rows pulled by TMA bulk copies into shared memory."""
    import ctypes
    lib = _lib.load()
    q = _lib.SpmmDesc()
    q.indptr, q.indices, q.indptr64 = ptr(indptr), ptr(indices), int(indptr.dtype == torch.int64)
    q.n_dst, q.n_src, q.d = indptr.numel() - 1, xq.data.shape[0], xq.cols
    q.X_q24, q.ldq = ptr(xq.data), xq.ldq
    q.Y_hi, q.Y_lo, q.ldyp = ptr(out_planes.hi), ptr(out_planes.lo), out_planes.hi.stride(0)
    q.self_add, q.mean_plus_one = int(self_add), int(mean_plus_one)
    check(lib.glnn_exp_spmm_tma_q24(ctypes.byref(q), int(stages), stream()), "glnn_exp_spmm_tma_q24")
    return out_planes


class S24:
    """EXPERIMENTAL sparse copy of a post-ReLU Q24 matrix (include/glnn_b200.h, glnn_compact_s24):
    .data int32 [rows, lds] words [fp32 bits 31..8 | column], .cap int32 [1] = max non-zeros per row
    (device)."""

    def __init__(self, data, cap, cols):
        self.data, self.cap, self.cols = data, cap, cols


def compact_s24(xq):
    """glnn_compact_s24: Q24 -> S24 (d <= 256)."""
    lib = _lib.load()
    require_cuda(xq.data)
    rows = xq.data.shape[0]
    lds = int(lib.glnn_s24_row_words(xq.cols))
    data = torch.empty(rows, lds, dtype=torch.int32, device=xq.data.device)
    cap = torch.zeros(1, dtype=torch.int32, device=xq.data.device)
    check(lib.glnn_compact_s24(ptr(xq.data), xq.ldq, rows, xq.cols, ptr(data), lds, ptr(cap), stream()),
          "glnn_compact_s24")
    return S24(data, cap, xq.cols)


def new_planes(rows, cols, device):
    ldp = (cols + 7) // 8 * 8
    return Planes(torch.zeros(rows, ldp, dtype=torch.int16, device=device),
                  torch.zeros(rows, ldp, dtype=torch.int16, device=device), cols)


def spmm_csr_q24_planes(indptr, indices, xq, self_add=False, mean_plus_one=False, src_scale=None,
                        dst_scale=None, out=None, hot_below=0):
    """Aggregation of a Q24 matrix into Planes (glnn_spmm_csr with X_q24 in, planes out)."""
    if out is None:
        out = new_planes(indptr.numel() - 1, xq.cols, xq.data.device)
    return spmm(indptr, indices, xq, out_planes=out, self_add=self_add, mean_plus_one=mean_plus_one,
                src_scale=src_scale, dst_scale=dst_scale, hot_below=hot_below)
