"""Shared helpers for the parity tests: load tests/golden/*.npz into the oracle's containers."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

TEACHER_CASES = ["sage_bn3", "sage_none2", "sage_wide", "gcn_cora_like", "gcn_agg_first"]
STUDENT_CASES = ["mlp_bn3", "mlp_lamb0", "mlp_none2_wd", "mlp_dropout", "mlp_small_n", "mlp_1layer"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def sub(d, prefix, dtype=None):
    """{'encoder.x': tensor} for every key starting with `prefix`, prefix and 'encoder.' stripped."""
    out = {}
    for k, v in d.items():
        if k.startswith(prefix):
            kk = k[len(prefix):]
            if kk.startswith("encoder."):
                kk = kk[len("encoder."):]
            t = torch.from_numpy(np.array(v))
            if dtype is not None and t.is_floating_point():
                t = t.to(dtype)
            out[kk] = t.clone()
    return out


def teacher_params(d, dtype=torch.float32):
    sd = sub(d, "sd.", dtype)
    L = int(d["num_layers"])
    sage = str(d["model_name"]) == "SAGE"
    layers, norms = [], []
    for l in range(L):
        if sage:
            layers.append((sd[f"layers.{l}.fc_neigh.weight"], sd[f"layers.{l}.fc_neigh.bias"]))
        else:
            layers.append((sd[f"layers.{l}.weight"], sd[f"layers.{l}.bias"]))
        if str(d["norm"]) == "batch" and l != L - 1:
            norms.append((sd[f"norms.{l}.weight"], sd[f"norms.{l}.bias"],
                          sd[f"norms.{l}.running_mean"], sd[f"norms.{l}.running_var"]))
    return layers, norms


def student_masks(d, idx_perm_count, num_layers):
    """Recorded keep-masks, regrouped as masks[pass][step][layer]."""
    n_masks = sum(1 for k in d if k.startswith("maskshape."))
    flat = []
    for i in range(n_masks):
        shape = tuple(int(x) for x in d[f"maskshape.{i}"])
        bits = np.unpackbits(d[f"mask.{i}"])[: shape[0] * shape[1]].reshape(shape)
        flat.append(torch.from_numpy(bits.astype(np.uint8)))
    return flat


def relerr_q(a, b, q=0.999):
    """q-quantile of |a-b| over max|b|: robust to the handful of Adam-amplified outliers (elements
    whose gradient is at the rounding-noise level move by O(lr) under ANY change of summation
    order, because the first Adam steps apply lr * g / (|g| + 1e-8))."""
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    d = (a - b).abs()
    kth = max(1, int(round(q * d.numel())))
    return float(d.kthvalue(kth).values / b.abs().max().clamp_min(1e-30))


def relerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class TorchKernels:
    """CPU test double for glnn_b200.ops (same keyword surface), built on the oracle's SpMM.  Used
    only to exercise HOST logic (sharding, relabelling, exchange plan) under gloo without a GPU."""

    @staticmethod
    def _epi(y, bias, col_scale, col_shift, relu, out):
        if bias is not None:
            y = y + bias
        if relu == 2:
            y = y.clamp(min=0)
        if col_scale is not None:
            y = y * col_scale + col_shift
        if relu == 1:
            y = y.clamp(min=0)
        if out is not None:
            out.copy_(y)
            return out
        return y

    @staticmethod
    def spmm_csr(indptr, indices, x, d=None, out=None, self_add=False, mean_plus_one=False,
                 src_scale=None, dst_scale=None, bias=None, col_scale=None, col_shift=None, relu=0):
        import glnn_oracle as O
        d = x.shape[1] if d is None else d
        xs = x[:, :d].contiguous()
        if src_scale is not None:
            xs = xs * src_scale.unsqueeze(1)
        y = O.spmm_sum(indptr.long(), indices.long(), xs, n_src=x.shape[0])
        n_dst = indptr.numel() - 1
        if self_add:
            y = y + x[:n_dst, :d]
        if mean_plus_one:
            deg = (indptr[1:] - indptr[:-1]).to(y.dtype).unsqueeze(1)
            y = y / (deg + 1)
        if dst_scale is not None:
            y = y * dst_scale.unsqueeze(1)
        return TorchKernels._epi(y, bias, col_scale, col_shift, relu, out)

    @staticmethod
    def gemm(a, b, trans_a=False, trans_b=False, out=None, row_scale=None, bias=None, col_scale=None,
             col_shift=None, relu=0, impl=0):
        y = (a.t() if trans_a else a) @ (b.t() if trans_b else b)
        if row_scale is not None:
            y = y * row_scale.unsqueeze(1)
        return TorchKernels._epi(y, bias, col_scale, col_shift, relu, out)

    @staticmethod
    def log_softmax(x, out=None):
        y = torch.log_softmax(x, 1)
        if out is not None:
            out.copy_(y)
            return out
        return y
