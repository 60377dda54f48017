"""Destination-row sharded SAGE teacher forward over the GPUs of one box.

The reference is single-process; this is how the same forward spreads over NVLink-connected B200s:
  * rows (destination nodes) are cut into G contiguous ranges balanced by NNZ (power-law degrees
    make node-count balance wrong); rank g owns CSR rows [cut[g], cut[g+1]);
  * every rank keeps a full replica of the current layer's embeddings in a PADDED layout
    [G * rows_max, d]: rank g's rows live at [g * rows_max, g * rows_max + rows_g).  Column ids of
    the local CSR slice are relabelled into that layout once, so the gather kernel indexes the
    replica directly and the exchange is one equal-sized all-gather per layer;
  * per layer: local aggregation + projection for the owned rows (same kernels as 1 GPU) and ONE
    all-gather of a [rows_max, d] slab (NCCL over NVLink 5 / NVSwitch) -- of the layer input for
    an aggregate-first layer, of the narrow projection for a project-first layer.  The final
    log-probabilities are gathered the same way.
"""
import torch
import torch.distributed as dist

from . import ops


ROW_COST = 16  # cost of one row (projection + output write + exchange) in units of one gathered edge


def nnz_balanced_cuts(indptr, world, row_cost=ROW_COST):
    """Row boundaries cut[0..world] with ~equal (nnz + row_cost * rows): the gather scales with the
    edges, the projection / stores / all-gather slab with the rows (measured on products: ~0.15 ns
    per edge at d=256 vs ~3 ns per row), so balancing edges alone leaves the low-degree shards with
    2.4x the rows of the hub shard and inflates the padded slab."""
    p = indptr.to(torch.int64).cpu()
    n = p.numel() - 1
    weight = p + row_cost * torch.arange(n + 1, dtype=torch.int64)  # monotone
    total = int(weight[-1])
    targets = torch.tensor([total * g // world for g in range(1, world)], dtype=torch.int64)
    inner = torch.searchsorted(weight, targets).clamp_(0, n)
    cuts = [0] + inner.tolist() + [n]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


class ShardedGraph:
    """Rank-local slice of a CSRGraph in the padded global layout."""

    def __init__(self, g, rank, world):
        self.rank, self.world, self.n = rank, world, g.num_nodes()
        self.cuts = nnz_balanced_cuts(g.indptr, world)
        self.rows_max = max(self.cuts[i + 1] - self.cuts[i] for i in range(world))
        r0, r1 = self.cuts[rank], self.cuts[rank + 1]
        self.r0, self.rows = r0, r1 - r0
        p = g.indptr.to(torch.int64)
        e0, e1 = int(p[r0]), int(p[r1])
        dev = g.indices.device
        cuts_t = torch.tensor(self.cuts, dtype=torch.int64, device=dev)
        cols = g.indices[e0:e1].to(torch.int64)
        owner = torch.searchsorted(cuts_t, cols, right=True) - 1
        cols = owner * self.rows_max + (cols - cuts_t[owner])
        local_ptr = p[r0:r1 + 1] - e0
        deg = local_ptr[1:] - local_ptr[:-1]
        # the SAGE "gcn" self term becomes an explicit edge to the row's own slot in the replica
        # (it sits at an offset there), and the mean uses 1 / (true in-degree + 1) as a row scale
        rows = torch.arange(self.rows, device=dev, dtype=torch.int64)
        self_pos = local_ptr[1:] + rows            # position of the appended self edge of each row
        nnz = int(local_ptr[-1]) + self.rows
        merged = torch.empty(nnz, dtype=torch.int64, device=dev)
        is_self = torch.zeros(nnz, dtype=torch.bool, device=dev)
        is_self[self_pos] = True
        merged[is_self] = rank * self.rows_max + rows
        merged[~is_self] = cols
        self.indices = merged.to(torch.int32)
        new_ptr = local_ptr + torch.arange(self.rows + 1, device=dev, dtype=torch.int64)
        self.indptr = new_ptr.to(torch.int32) if nnz < 2 ** 31 else new_ptr
        self.inv_deg1 = (1.0 / (deg.to(torch.float32) + 1.0)).contiguous()

    def to_padded(self, x):
        """[n, d] in original node order -> [world * rows_max, d] padded layout (zeros in pads)."""
        out = torch.zeros(self.world * self.rows_max, x.shape[1], dtype=x.dtype, device=x.device)
        for g in range(self.world):
            a, b = self.cuts[g], self.cuts[g + 1]
            out[g * self.rows_max: g * self.rows_max + (b - a)] = x[a:b]
        return out

    def from_padded(self, xp):
        return torch.cat([xp[g * self.rows_max: g * self.rows_max + (self.cuts[g + 1] - self.cuts[g])]
                          for g in range(self.world)])


def _pad_rows(w, b, dpad):
    """Zero-pads an nn.Linear weight [d_out, d_in] / bias [d_out] to dpad output rows."""
    d_out = w.shape[0]
    if dpad == d_out:
        return w, b
    wp = torch.zeros(dpad, w.shape[1], dtype=w.dtype, device=w.device)
    wp[:d_out] = w
    bp = torch.zeros(dpad, dtype=b.dtype, device=b.device)
    bp[:d_out] = b
    return wp, bp


def _buffer(sg, key, rows, cols, like):
    """Per-shard cache of activation buffers: a forward reuses the same HBM every call (no
    allocator traffic, no memsets of multi-GB tensors inside the timed region).  Pad rows are never
    read by the gathers (no column id points at them), so their content is irrelevant."""
    cache = sg.__dict__.setdefault("_bufs", {})
    t = cache.get(key)
    if t is None or t.shape != (rows, cols) or t.device != like.device or t.dtype != like.dtype:
        t = torch.zeros(rows, cols, dtype=like.dtype, device=like.device)
        cache[key] = t
    return t


def sage_forward_sharded(sg, feats_pad, layers, norms, group=None, kernels=None, log_softmax=True,
                         timings=None, gather_output=True):
    """layers: [(W [out,in], b)], norms: [(scale, shift)] folded eval-BN per hidden layer (or []).
    feats_pad: padded replica of the input features (valid on every rank).  Returns the padded
    replica [world * rows_max, label_dim] of the output (log-probabilities when log_softmax); the
    returned tensor is a cached buffer that the next call overwrites.  With gather_output=False
    the result stays sharded: only this rank's slab [rank*rows_max, +rows) of it is valid (what a
    sharded consumer -- loss/accuracy reduction, a sharded student -- needs).

    Exchange plan: an aggregate-first layer needs the full replica of its input (all-gather of the
    previous output, d_in wide); a project-first layer (4-padded d_out < d_in) projects only the
    owned rows and all-gathers the narrow projection instead -- for ogbn-products the last layer
    moves 48 instead of 256 columns.  `kernels` (default: the CUDA ops) is injectable so the host
    logic can be exercised with gloo on CPU by the tests."""
    k = kernels or ops
    world, rm = sg.world, sg.rows_max
    lo, hi = sg.rank * rm, sg.rank * rm + sg.rows

    def mark(name):
        if timings is not None and feats_pad.is_cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timings.append((name, ev))

    def gather(buf):
        if world > 1:  # in place: this rank's slab already sits at its offset in the output
            dist.all_gather_into_tensor(buf, buf[lo: lo + rm], group=group)
        mark(f"all_gather {tuple(buf.shape)} {buf.dtype}".replace("torch.", ""))

    mark("start")

    h, h_full = feats_pad, True   # h: fp32 replica, or ops.Q24 replica after a q24 hand-off
    L = len(layers)
    planes_ok = hasattr(k, "spmm_csr_planes")

    def proj_first(i):
        d_o, d_i = layers[i][0].shape
        sc = norms[i][0] if (norms and i != L - 1) else None
        return ((d_o + 3) // 4 * 4) < d_i and (sc is None or d_o % 4 == 0)

    for l, (w, b) in enumerate(layers):
        last = l == L - 1
        scale, shift = (norms[l] if (norms and not last) else (None, None))
        relu = 0 if last else 1
        d_out, d_in = w.shape
        dpad = (d_out + 3) // 4 * 4
        if proj_first(l):
            wp, bp = _pad_rows(w, b, dpad)
            z = _buffer(sg, ("z", l), world * rm, dpad, feats_pad)
            k.gemm(h[lo:hi, :d_in], wp, trans_b=True, out=z[lo:hi])
            mark(f"L{l} gemm {d_in}->{dpad}")
            gather(z)
            y = _buffer(sg, ("y", l), world * rm, dpad, feats_pad)
            k.spmm_csr(sg.indptr, sg.indices, z, d=dpad, out=y[lo:hi], dst_scale=sg.inv_deg1,
                       bias=bp, col_scale=scale, col_shift=shift, relu=relu)
            mark(f"L{l} spmm d={dpad}")
        else:
            is_q24 = planes_ok and isinstance(h, k.Q24)
            if not h_full:
                gather(h.data if is_q24 else h)
            # the output feeds another gather and nothing else -> exchange it as 24-bit rows
            out_q24 = planes_ok and not last and not proj_first(l + 1) and d_out % 8 == 0 \
                and d_out <= 512
            if planes_ok:
                # gather straight into bf16 hi/lo planes, the tensor-core projection's operand format
                ldp = (d_in + 7) // 8 * 8
                pl = sg.__dict__.setdefault("_planes", {}).get(l)
                if pl is None or pl.hi.shape != (max(sg.rows, 1), ldp):
                    mk = lambda: torch.zeros(max(sg.rows, 1), ldp, dtype=torch.int16,
                                             device=feats_pad.device)
                    pl = k.Planes(mk(), mk(), d_in)
                    sg._planes[l] = pl
                if is_q24:
                    k.spmm_csr_q24_planes(sg.indptr, sg.indices, h, dst_scale=sg.inv_deg1, out=pl)
                else:
                    k.spmm_csr_planes(sg.indptr, sg.indices, h, d=d_in, dst_scale=sg.inv_deg1, out=pl)
                mark(f"L{l} spmm d={d_in}")
                wpl = k.split_planes(w)
                if out_q24:
                    cache = sg.__dict__.setdefault("_bufs", {})
                    yq = cache.get(("yq", l))
                    ldq = k.Q24.row_bytes(d_out)
                    if yq is None or yq.shape != (world * rm, ldq):
                        yq = torch.zeros(world * rm, ldq, dtype=torch.uint8, device=feats_pad.device)
                        cache[("yq", l)] = yq
                    k.gemm_planes_q24(pl, wpl, out=k.Q24(yq[lo:hi], d_out), bias=b, col_scale=scale,
                                      col_shift=shift, relu=relu)
                    y = k.Q24(yq, d_out)
                else:
                    y = _buffer(sg, ("y", l), world * rm, dpad, feats_pad)
                    k.gemm_planes(pl, wpl, trans_b=True, out=y[lo:hi, :d_out], bias=b, col_scale=scale,
                                  col_shift=shift, relu=relu)
            else:
                y = _buffer(sg, ("y", l), world * rm, dpad, h)
                agg = _buffer(sg, ("agg", l), max(sg.rows, 1), (d_in + 3) // 4 * 4, h)
                k.spmm_csr(sg.indptr, sg.indices, h, d=d_in, out=agg[: sg.rows, :d_in],
                           dst_scale=sg.inv_deg1)
                mark(f"L{l} spmm d={d_in}")
                k.gemm(agg[: sg.rows, :d_in], w, trans_b=True, out=y[lo:hi, :d_out], bias=b,
                       col_scale=scale, col_shift=shift, relu=relu)
            mark(f"L{l} gemm {d_in}->{d_out}")
        h, h_full = y, False
    c = layers[-1][0].shape[0]
    out = _buffer(sg, ("out",), world * rm, c, h)
    if log_softmax:
        k.log_softmax(h[lo:hi, :c], out=out[lo:hi])
    else:
        out[lo:hi] = h[lo:hi, :c]
    mark("log_softmax")
    if gather_output:  # otherwise every rank keeps only its own slab valid (rows lo:hi)
        gather(out)
    return out
