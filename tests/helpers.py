"""Shared helpers for the parity tests: load tests/golden/*.npz into the oracle's containers."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

TEACHER_CASES = ["sage_bn3", "sage_none2", "sage_wide", "gcn_cora_like", "gcn_agg_first", "gcn_1layer"]
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


def parity_report(a, b, rtol=1e-4, atol=1e-5):
    """The two halves of SURVEY.md 8d's parity gate for an output tensor: max|a-b| / max|b| and
    allclose(rtol, atol) elementwise, plus how far the worst element is from its allclose bound."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    diff = (a - b).abs()
    bound = atol + rtol * b.abs()
    viol = diff > bound
    return {"max_rel": float(diff.max() / b.abs().max().clamp_min(1e-30)),
            "allclose": not bool(viol.any()), "violating_frac": float(viol.double().mean()),
            "worst_over_bound": float((diff / bound).max())}


def assert_parity(a, b, what="", tol=1e-4, rtol=1e-4, atol=1e-5):
    """SURVEY.md 8d gate: max|a-b| / max|b| <= 1e-4 AND allclose(rtol=1e-4, atol=1e-5)."""
    r = parity_report(a, b, rtol, atol)
    assert r["max_rel"] <= tol and r["allclose"], (what, r)
    return r


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
        if int(indptr[0]) != 0:  # a row range of a larger CSR (absolute offsets, as the kernel takes it)
            indices = indices[int(indptr[0]): int(indptr[-1])]
            indptr = indptr - indptr[0]
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


def dp_student_pass(dist, rank, world, p, state, feats, targets, kind, idx_batch, lamb, num_layers,
                    norm, lr, weight_decay, eps=1e-5, momentum=0.1):
    """Test double of glnn_mlp_train_pass_dp (csrc/mlp.cu) in plain torch over any process group:
    the SAME decomposition the kernels use, so that the protocol can be checked on CPU (gloo):
      * rank r takes rows [r, r+1) * B / world of every global batch;
      * BatchNorm: per-rank (mean, M2) -> all_gather -> Chan combination in rank order -> every rank
        normalises with the statistics of the GLOBAL batch and updates its running stats with them;
      * backward: per-rank S1 = sum g, S2 = sum g * xhat -> all_gather -> sums; dgamma / dbeta are
        already global, so only rank 0 contributes them to the gradient reduction;
      * d(lamb * loss)/dlogits is scaled by 1 / B_global;
      * optimizer: rank r sums slice r of the flat gradient over the ranks (rank order), applies Adam
        to that slice with its slice of the moments, the new parameter slices are all-gathered.
    p / state: the oracle's dicts (glnn_oracle.mlp_forward), updated in place.  Returns the pass's
    mean unscaled loss (sum of the ranks' shares)."""
    import math
    import torch
    keys = [k for k in p if not (k.endswith("running_mean") or k.endswith("running_var")
                                 or k.endswith("num_batches_tracked"))]
    nb, B = idx_batch.shape
    R = B // world

    def gather_cat(t):
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t.contiguous())
        return outs

    total = 0.0
    for i in range(nb):
        rows = idx_batch[i, rank * R:(rank + 1) * R]
        h = feats[rows]
        cache = []
        for l in range(num_layers):
            w, b = p[f"layers.{l}.weight"], p[f"layers.{l}.bias"]
            z = h @ w.t() + b
            if l == num_layers - 1:
                cache.append((h, None, None, None))
                h = z
                break
            if norm == "batch":
                mu_l = z.mean(0)
                m2_l = ((z - mu_l) ** 2).sum(0)
                parts = gather_cat(torch.stack([mu_l, m2_l]))
                n, mu, m2 = 0.0, torch.zeros_like(mu_l), torch.zeros_like(mu_l)
                for part in parts:                      # Chan, rank order
                    delta, tot = part[0] - mu, n + R
                    mu = mu + delta * (R / tot)
                    m2 = m2 + part[1] + delta * delta * (n * R / tot)
                    n = tot
                var = m2 / B
                p[f"norms.{l}.running_mean"].mul_(1 - momentum).add_(momentum * mu)
                p[f"norms.{l}.running_var"].mul_(1 - momentum).add_(momentum * m2 / max(B - 1, 1))
                p[f"norms.{l}.num_batches_tracked"] += 1
                invstd = 1.0 / torch.sqrt(var + eps)
                xhat = (z - mu) * invstd
                y = xhat * p[f"norms.{l}.weight"] + p[f"norms.{l}.bias"]
            else:
                xhat, invstd, y = None, None, z
            cache.append((h, xhat, invstd, y))
            h = torch.relu(y)
        logits = h
        s = torch.log_softmax(logits, dim=1)
        tg = targets[rows]
        if kind == "nll":
            loss_share = -s[torch.arange(R), tg].sum() / B
            d = s.exp()
            d[torch.arange(R), tg] -= 1.0
        else:
            et = tg.exp()
            loss_share = (et * (tg - s)).sum() / B
            d = s.exp() * et.sum(1, keepdim=True) - et
        d = d * (lamb / B)
        total += float(loss_share)
        grads = {}
        for l in reversed(range(num_layers)):
            h_in, xhat, invstd, y = cache[l]
            if l != num_layers - 1:
                d = d * (y > 0).to(d.dtype)
                if norm == "batch":
                    parts = gather_cat(torch.stack([d.sum(0), (d * xhat).sum(0)]))
                    S1, S2 = sum(q[0] for q in parts), sum(q[1] for q in parts)
                    grads[f"norms.{l}.bias"] = S1 if rank == 0 else torch.zeros_like(S1)
                    grads[f"norms.{l}.weight"] = S2 if rank == 0 else torch.zeros_like(S2)
                    d = (p[f"norms.{l}.weight"] * invstd / B) * (B * d - S1 - xhat * S2)
            grads[f"layers.{l}.weight"] = d.t() @ h_in
            grads[f"layers.{l}.bias"] = d.sum(0)
            if l > 0:
                d = d @ p[f"layers.{l}.weight"]
        # fused optimizer: flat buffers, equal slices of a multiple of 4 elements
        flat_g = torch.cat([grads[k].reshape(-1) for k in keys])
        P = flat_g.numel()
        sl = (P + 4 * world - 1) // (4 * world) * 4
        pad = lambda t: torch.cat([t, t.new_zeros(sl * world - P)])
        flat_g = pad(flat_g)
        flat_p = pad(torch.cat([p[k].reshape(-1) for k in keys]))
        flat_m = pad(torch.cat([state[k]["exp_avg"].reshape(-1) for k in keys]))
        flat_v = pad(torch.cat([state[k]["exp_avg_sq"].reshape(-1) for k in keys]))
        lo, hi = rank * sl, (rank + 1) * sl
        g = sum(q[lo:hi] for q in gather_cat(flat_g))      # P2P loads of slice `rank`, rank order
        t = state[keys[0]]["step"] + 1
        if weight_decay != 0:
            g = g + weight_decay * flat_p[lo:hi]
        m = flat_m[lo:hi] * 0.9 + 0.1 * g
        v = flat_v[lo:hi] * 0.999 + 0.001 * g * g
        denom = v.sqrt() / math.sqrt(1 - 0.999 ** t) + 1e-8
        new_p = flat_p[lo:hi] - (lr / (1 - 0.9 ** t)) * m / denom
        all_p, all_m, all_v = (torch.cat(gather_cat(x)) for x in (new_p, m, v))
        off = 0
        for k in keys:
            n = p[k].numel()
            p[k].copy_(all_p[off:off + n].view_as(p[k]))
            state[k]["exp_avg"].copy_(all_m[off:off + n].view_as(p[k]))
            state[k]["exp_avg_sq"].copy_(all_v[off:off + n].view_as(p[k]))
            state[k]["step"] = t
            off += n
    tot = torch.tensor([total], dtype=torch.float64)
    dist.all_reduce(tot)
    return float(tot) / nb
