"""Destination-row sharded SAGE teacher forward over the GPUs of one box.

The reference is single-process; this is how the same forward spreads over NVLink-connected B200s:
  * rows (destination nodes) are cut into G contiguous ranges balanced by nnz + row cost (power-law
    degrees make node-count balance wrong); rank g owns CSR rows [cut[g], cut[g+1]);
  * every rank keeps a full replica of the matrix the next gather reads, in a PADDED, CHUNKED layout
    [C chunks][G ranks][rc rows]: local row r of rank g lives at (r // rc) * G * rc + g * rc + r % rc.
    Column ids of the local CSR slice are relabelled into that layout once, so the gather kernel
    indexes the replica directly, and chunk c of every rank forms one contiguous block that a single
    in-place all-gather fills;
  * per layer: local aggregation + projection of the owned rows (same kernels as 1 GPU: q24 gathers,
    tcgen05 projections) and ONE exchange of the matrix the next gather needs -- the layer output
    for an aggregate-first successor, the narrow projection for a project-first layer (ogbn-products'
    last layer moves 48 instead of 256 columns).  The exchange is pipelined: the rows are processed
    in C chunks and the NCCL all-gather of chunk c (side stream, NVLink 5 / NVSwitch) overlaps the
    aggregation + projection of chunk c + 1;
  * what is exchanged is the 24-bit row-packed format the gather reads anyway (768 instead of 1024
    bytes per 256-wide row).
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
    """Rank-local slice of a CSRGraph in the padded, chunked global layout."""

    def __init__(self, g, rank, world, chunks=None):
        self.rank, self.world, self.n = rank, world, g.num_nodes()
        self.cuts = nnz_balanced_cuts(g.indptr, world)
        rows_max = max(max(self.cuts[i + 1] - self.cuts[i] for i in range(world)), 1)
        if chunks is None:  # pipeline the exchange only when a chunk is still a big kernel
            import os
            want = int(os.environ.get("GLNN_DIST_CHUNKS", "4"))
            chunks = want if (world > 1 and rows_max >= want * 32768) else 1
        self.chunks = chunks
        self.rc = (rows_max + chunks - 1) // chunks          # rows per chunk and rank
        self.rows_max = self.rc * chunks                      # padded rows per rank
        self.total_rows = world * self.rows_max
        r0, r1 = self.cuts[rank], self.cuts[rank + 1]
        self.r0, self.rows = r0, r1 - r0
        p = g.indptr.to(torch.int64)
        e0, e1 = int(p[r0]), int(p[r1])
        dev = g.indices.device
        cuts_t = torch.tensor(self.cuts, dtype=torch.int64, device=dev)
        self._cuts_t = cuts_t
        cols = self._pad_ids(g.indices[e0:e1].to(torch.int64))
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
        merged[is_self] = self._pad_ids(rows + r0)
        merged[~is_self] = cols
        self.indices = merged.to(torch.int32)
        new_ptr = local_ptr + torch.arange(self.rows + 1, device=dev, dtype=torch.int64)
        self.indptr = new_ptr.to(torch.int32) if nnz < 2 ** 31 else new_ptr
        self.inv_deg1 = (1.0 / (deg.to(torch.float32) + 1.0)).contiguous()
        self.pad_ids = self._pad_ids(torch.arange(self.n, device=dev, dtype=torch.int64))

    def _pad_ids(self, nodes):
        """Original node ids -> row in the padded, chunked replica layout."""
        owner = torch.searchsorted(self._cuts_t, nodes, right=True) - 1
        owner.clamp_(0, self.world - 1)
        local = nodes - self._cuts_t[owner]
        return (local // self.rc) * (self.world * self.rc) + owner * self.rc + local % self.rc

    def early_owners(self):
        """Source owners whose slabs a rank waits for FIRST when the consuming aggregation is split
        into two passes: itself and the next world/2 - 1 ranks (so that every rank is "early" for
        world/2 - 1 peers and "late" for the others)."""
        h = max(1, self.world // 2)
        return [(self.rank + i) % self.world for i in range(h)]

    def early_receivers(self):
        """Peers that need THIS rank's slab first (those that list it among their early owners)."""
        h = max(1, self.world // 2)
        return [(self.rank - i) % self.world for i in range(1, h)]

    def split_by_owner(self):
        """The local CSR cut into two CSRs over the same rows: edges whose source row is owned by an
        early owner (the self edge included) and the rest.  Cached."""
        hit = self.__dict__.get("_split")
        if hit is None:
            dev = self.indices.device
            cols = self.indices.to(torch.int64)
            owner = (cols // self.rc) % self.world
            early_mask = torch.zeros(self.world, dtype=torch.bool, device=dev)
            early_mask[torch.tensor(self.early_owners(), device=dev)] = True
            is_early = early_mask[owner]
            ptr = self.indptr.to(torch.int64)
            deg = ptr[1:] - ptr[:-1]
            row = torch.repeat_interleave(torch.arange(self.rows, device=dev), deg)
            out = []
            for m in (is_early, ~is_early):
                cnt = torch.bincount(row[m], minlength=self.rows)
                p = torch.zeros(self.rows + 1, dtype=torch.int64, device=dev)
                torch.cumsum(cnt, 0, out=p[1:])
                idx = self.indices[m].contiguous()          # boolean mask keeps the CSR order
                out.append((p.to(torch.int32) if idx.numel() < 2 ** 31 else p, idx))
            hit = tuple(out)
            self._split = hit
        return hit

    def chunk_rows(self, c):
        """Local row range [a, b) of chunk c (may be empty for the last chunks of a short shard)."""
        a = min(self.rows, c * self.rc)
        return a, min(self.rows, a + self.rc)

    def slab_start(self, c, rank=None):
        """First replica row of chunk c of `rank` (default: this rank)."""
        return c * self.world * self.rc + (self.rank if rank is None else rank) * self.rc

    def to_padded(self, x):
        """[n, d] in original node order -> [total_rows, d] padded replica (zeros in the pads)."""
        out = torch.zeros(self.total_rows, x.shape[1], dtype=x.dtype, device=x.device)
        out[self.pad_ids.to(x.device)] = x
        return out

    def from_padded(self, xp):
        return xp[self.pad_ids.to(xp.device)]

    def local_rows_of(self, xp):
        """This rank's rows [rows, d] (local order) out of a padded replica."""
        return xp[self.pad_ids[self.r0:self.r0 + self.rows].to(xp.device)]


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


def _cached(sg, key, make):
    """Per-shard cache of activation buffers: a forward reuses the same HBM every call (no allocator
    traffic, no memsets of multi-GB tensors inside the timed region).  Pad rows are never read by
    the gathers (no column id points at them), so their content is irrelevant."""
    cache = sg.__dict__.setdefault("_bufs", {})
    t = cache.get(key)
    if t is None:
        t = make()
        cache[key] = t
    return t


def _push_mode():
    """Exchange transport on CUDA: "push" (default) = every rank copies its slab straight into the
    peers' replicas through peer-mapped symmetric memory (copy engines over NVLink: no SMs, no
    staging, so it overlaps the aggregation kernels that fill every SM); "nccl" = all-gather."""
    import os
    return os.environ.get("GLNN_EXCHANGE", "push").lower() != "nccl"


def _push_engine():
    """Who moves a slab to the peers: "sm" (default) = glnn_peer_push, a small kernel whose posted
    stores run the links at ~0.75 TB/s per direction; "ce" = one cudaMemcpyPeerAsync per peer on the
    copy engines (round 1: 0.37-0.45 TB/s per rank at N=8 under the HBM-saturating gather)."""
    import os
    return os.environ.get("GLNN_PUSH_ENGINE", "sm").lower()


def _push_ctas():
    import os
    return int(os.environ.get("GLNN_PUSH_CTAS", "32"))


def _two_pass(world=8, which="TWO_PASS"):
    """Two-pass consumption of an exchanged replica (GLNN_DIST_TWO_PASS = 0 / 1; default: on from 4
    ranks up -- measured on ogbn-products: 8 ranks 8.24 -> 7.68 ms, 2 ranks 17.25 -> 18.15 ms, where
    the two half-passes cost more than the exchange they hide): the producing
    layer pushes every chunk to the peers that need it first, then to the others; the consuming
    aggregation runs a first pass over the edges whose sources are already here (own rows + early
    owners: half of the edges) while the late slabs are still in flight, and a second pass over the
    rest on top of the fp32 partial sums (glnn_spmm_csr Y_init)."""
    import os
    v = os.environ.get("GLNN_DIST_" + which, os.environ.get("GLNN_DIST_TWO_PASS", ""))
    if v == "":
        return world >= 4
    return v != "0"


def _replicate_projection():
    """Exchange policy for an aggregate-first layer whose successor is aggregate-first as well
    (ogbn-products layer 0: 100 -> 256).  "1" (opt-in, GLNN_DIST_REPLICATE): ship the NARROW aggregated operand (d_in
    columns as bf16 hi/lo planes) and let every rank project all N rows itself -- the projection is
    HBM-bound and cheap (0.5 ms for all 2.45 M rows), the exchange is what limits scaling (round 1:
    3.9 of 9.75 ms exposed at N=8 for 1.65 GB per rank of 256-wide q24 rows; the planes of the
    100-wide operand are 0.89 GB).  "0" (default): project the owned rows and ship the wide q24
    output.  Measured at N=2: replication costs 1.38 ms for the all-rows projection and loses 0.6 ms
    per forward; it only pays where the exchange cannot be hidden."""
    import os
    return os.environ.get("GLNN_DIST_REPLICATE", "0") not in ("", "0")


def _symm_replica(sg, key, rows, row_bytes, dev, group):
    """uint8 [rows, row_bytes] replica allocated in symmetric memory (cached per shard): returns
    (tensor, handle).  Collective: every rank reaches the same allocation in the same order."""
    cache = sg.__dict__.setdefault("_bufs", {})
    hit = cache.get(("symm",) + key)
    if hit is None:
        import torch.distributed._symmetric_memory as symm
        flat = symm.empty(rows * row_bytes, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(flat, group or dist.group.WORLD)
        flat.zero_()
        data = flat.view(rows, row_bytes)
        peers = {r: hdl.get_buffer(r, (rows, row_bytes), torch.uint8) for r in range(sg.world)
                 if r != sg.rank}
        torch.cuda.synchronize(dev)
        hdl.barrier(channel=0)   # nobody pushes into a replica that is still being zeroed
        hit = (data, hdl, peers)
        cache[("symm",) + key] = hit
        sg.__dict__.setdefault("_symm_by_ptr", {})[data.data_ptr()] = hit
    return hit


class _Exchange:
    """Chunk-wise in-place exchange of a replica buffer on a side stream, so that the exchange of
    chunk c overlaps the computation of chunk c + 1 (CUDA); synchronous on CPU / gloo.  Replicas that
    live in symmetric memory are exchanged by peer copies + one device-side barrier per layer, all
    other buffers by NCCL all-gather."""

    def __init__(self, sg, group, cuda, mark):
        self.sg, self.group, self.cuda, self.mark = sg, group, cuda, mark
        self.stream = None
        self.pending = None   # symmetric-memory handle whose pushes still need their barrier
        if cuda and sg.world > 1:
            # high-priority side streams: whatever the driver uses for the transfer (copy engine, a
            # copy kernel, NCCL's kernels) should not queue behind the aggregation grid that keeps
            # every SM full.  Measured at N=4: neither priority nor one stream per peer changes the
            # exposed exchange (~2 ms of a 1.4 GB-per-rank transfer): the transfer does not speed up
            # beyond ~700 GB/s per rank while the HBM-saturating gather runs (profiles/r1i_bench_n4.json)
            self.stream = _cached(sg, ("comm_stream",), lambda: torch.cuda.Stream(priority=-1))
            # one stream per peer: the G-1 peer copies of a chunk run concurrently
            self.peer_streams = _cached(sg, ("peer_streams",),
                                        lambda: [torch.cuda.Stream(priority=-1)
                                                 for _ in range(sg.world - 1)])

    def chunk_two_phase(self, buf2d, c, events):
        """Two-phase variant of chunk(): records the chunk's ready event; the pushes themselves are
        issued by flush_two_phase() -- first every chunk to the EARLY receivers, then to the rest."""
        ev = torch.cuda.Event()
        ev.record()
        events.append((c, ev))

    def flush_two_phase(self, buf2d, events):
        """Issues, on the push stream: [chunk -> early receivers] for all chunks, a barrier across the
        ranks ("early slabs have landed everywhere"), [chunk -> late receivers] for all chunks, a second
        barrier.  Returns (early_done, late_done) events for the consumer to wait on."""
        sg = self.sg
        _, hdl, peers = sg.__dict__["_symm_by_ptr"][buf2d.data_ptr()]
        early = sg.early_receivers()
        late = [(sg.rank + i) % sg.world for i in range(1, sg.world) if (sg.rank + i) % sg.world not in early]
        st = self.peer_streams[0]
        done = []
        for group, channel in ((early, 2), (late, 3)):
            for c, ev in events:
                s0 = sg.slab_start(c)
                st.wait_event(ev)
                if group:
                    with torch.cuda.stream(st):
                        ops.peer_push(buf2d[s0: s0 + sg.rc], [peers[p][s0: s0 + sg.rc] for p in group],
                                      _push_ctas())
            with torch.cuda.stream(st):
                hdl.barrier(channel=channel)
                e = torch.cuda.Event()
                e.record()
            done.append(e)
        return done

    def chunk(self, buf2d, c):
        sg = self.sg
        if sg.world == 1:
            return
        lo = c * sg.world * sg.rc
        whole = buf2d[lo: lo + sg.world * sg.rc]
        s0 = sg.slab_start(c)
        mine = buf2d[s0: s0 + sg.rc]
        if self.stream is None:
            dist.all_gather_into_tensor(whole, mine, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()
        symm_hit = sg.__dict__.get("_symm_by_ptr", {}).get(buf2d.data_ptr())
        if symm_hit is None:
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(whole, mine, group=self.group)
            return
        _, hdl, peers = symm_hit
        if _push_engine() == "sm":
            # one SM-driven kernel: a 16-byte load of the slab feeds G-1 posted NVLink stores
            # (staggered targets: rank r writes r+1, r+2, ... first)
            order = [(sg.rank + i) % sg.world for i in range(1, sg.world)]
            st = self.peer_streams[0]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                ops.peer_push(mine, [peers[p][s0: s0 + sg.rc] for p in order], _push_ctas())
        else:
            for i in range(1, sg.world):   # copy engines: rank r's stream i writes to rank r+i
                peer = (sg.rank + i) % sg.world
                st = self.peer_streams[i - 1]
                st.wait_event(ev)
                with torch.cuda.stream(st):
                    peers[peer][s0: s0 + sg.rc].copy_(mine, non_blocking=True)
        self.pending = hdl

    def wait(self, what):
        if self.stream is not None:
            if self.pending is not None:   # every rank's pushes of this layer have landed
                for st in self.peer_streams:
                    self.stream.wait_stream(st)
                with torch.cuda.stream(self.stream):
                    self.pending.barrier(channel=0)
                self.pending = None
            torch.cuda.current_stream().wait_stream(self.stream)
        self.mark(what)


def sage_forward_sharded(sg, feats_pad, layers, norms, group=None, kernels=None, log_softmax=True,
                         timings=None, gather_output=True, feats_local=None, split_only=False):
    """layers: [(W [out,in], b)], norms: [(scale, shift)] folded eval-BN per hidden layer (or []).
    feats_pad: padded replica (sg.to_padded) of the input features, valid on every rank -- OR None
    with feats_local = this rank's OWN feature rows [sg.rows, d_in] (local order): the replica of the
    input is then assembled by the same chunked exchange as every other layer (each rank quantises
    its rows, NVLink all-gather), so a host only ships 1/G of the features to each GPU.  Returns
    the padded replica [total_rows, label_dim] of the output (log-probabilities when log_softmax);
    the returned tensor is a cached buffer that the next call overwrites.  With gather_output=False
    the result stays sharded: only this rank's rows of it are valid (sg.local_rows_of) -- what a
    sharded consumer (loss / accuracy reduction, a sharded student) needs.

    split_only=True: only the owner-split CSR (sg.split_by_owner()) is trusted -- sg.indptr / sg.indices
    are not read.  The host pipeline uses it with the two-pass exchange so that it uploads each edge
    once: aggregations whose input is complete run as two launches over the two halves.

    Exchange plan: an aggregate-first layer needs the full replica of its input (all-gather of the
    previous output, d_in wide); a project-first layer (4-padded d_out < d_in) projects only the
    owned rows and all-gathers the narrow projection instead.  `kernels` (default: the CUDA ops) is
    injectable so that the host logic -- cuts, relabelling, chunked layout, exchange plan -- can be
    exercised with gloo on CPU by the tests (fp32 torch double: same control flow, fp32 replicas)."""
    k = kernels or ops
    cuda = hasattr(k, "Q24")           # the real kernels: q24 replicas, planes operands
    world = sg.world
    if (feats_pad is None) == (feats_local is None):
        raise ValueError("sage_forward_sharded: give exactly one of feats_pad / feats_local")
    feats_any = feats_pad if feats_pad is not None else feats_local
    if feats_local is not None and feats_local.shape[0] != sg.rows:
        raise ValueError(f"feats_local has {feats_local.shape[0]} rows, this shard owns {sg.rows}")
    dev = feats_any.device
    L = len(layers)
    C = sg.chunks

    def mark(name):
        if timings is not None and feats_any.is_cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timings.append((name, ev))

    xch = _Exchange(sg, group, feats_any.is_cuda, mark)
    if cuda and world > 1 and _push_mode():
        # The replicas are cached and reused by the next call.  A fast rank must not push its slab of
        # call k+1 into a peer's replica while that peer is still gathering from it in call k (an
        # aggregate-first last layer with gather_output=False has no later synchronisation): one
        # device-side barrier over symmetric memory, ordered after everything this rank enqueued.
        hits = sg.__dict__.get("_symm_by_ptr")
        if hits:
            next(iter(hits.values()))[1].barrier(channel=1)
    mark("start")

    def proj_first(i):
        d_o, d_i = layers[i][0].shape
        sc = norms[i][0] if (norms and i != L - 1) else None
        return ((d_o + 3) // 4 * 4) < d_i and (sc is None or d_o % 4 == 0)

    # ---- format helpers: a "replica" is what a gather reads, an "operand" what a projection reads
    def new_replica(key, d):
        if cuda and world > 1 and _push_mode():
            data, _, _ = _symm_replica(sg, key, sg.total_rows, k.Q24.row_bytes(d), dev, group)
            return k.Q24(data, d)
        if cuda:
            return _cached(sg, key, lambda: k.Q24.empty(sg.total_rows, d, dev, zero=True))
        return _cached(sg, key, lambda: torch.zeros(sg.total_rows, d, dtype=torch.float32, device=dev))

    def replica_2d(rep):
        return rep.data if cuda else rep

    def new_operand(key, d):
        if cuda:
            return _cached(sg, key, lambda: k.new_planes(max(sg.rows, 1), d, dev))
        return _cached(sg, key, lambda: torch.zeros(max(sg.rows, 1), d, dtype=torch.float32, device=dev))

    def rows_of(op, a, b):
        if cuda:
            return k.Planes(op.hi[a:b], op.lo[a:b], op.cols)
        return op[a:b]

    def aggregate(rep, d, a, b, out_op):
        """mean over (neighbours + self) of replica rows for local rows [a, b) -> operand rows."""
        if cuda and split_only:
            (pe, ie), (pl_, il) = sg.split_by_owner()
            part = _cached(sg, ("partial_so", d), lambda: torch.empty(max(sg.rows, 1), (d + 7) // 8 * 8,
                                                                       dtype=torch.float32, device=dev))
            k.spmm(pe[a:b + 1], ie, rep, d=d, out=part[a:b])
            k.spmm(pl_[a:b + 1], il, rep, d=d, out_planes=rows_of(out_op, a, b),
                   dst_scale=sg.inv_deg1[a:b], acc_init=part[a:b])
        elif cuda:
            k.spmm(sg.indptr[a:b + 1], sg.indices, rep, d=d, out_planes=rows_of(out_op, a, b),
                   dst_scale=sg.inv_deg1[a:b])
        else:
            k.spmm_csr(sg.indptr[a:b + 1], sg.indices, rep, d=d, out=out_op[a:b],
                       dst_scale=sg.inv_deg1[a:b])

    def project(op_rows, w, wpl, b, scale, shift, relu, dest, a, b_row, c):
        """dest: ("replica", rep) -> rows of chunk c in the gather format; ("operand", op) -> local
        rows [a, b_row); ("final", out2d) -> fp32 rows of the padded output."""
        kind, buf = dest
        kw = dict(bias=b, col_scale=scale, col_shift=shift, relu=relu)
        if kind == "replica":
            s0 = sg.slab_start(c)
            if cuda:
                k.gemm_planes_q24(op_rows, wpl, out=k.Q24(buf.data[s0:s0 + (b_row - a)], buf.cols), **kw)
            else:
                k.gemm(op_rows, w, trans_b=True, out=buf[s0:s0 + (b_row - a)], **kw)
        elif kind == "operand":
            if cuda:
                k.gemm_planes(op_rows, wpl, trans_b=True, out_planes=rows_of(buf, a, b_row), **kw)
            else:
                k.gemm(op_rows, w, trans_b=True, out=buf[a:b_row], **kw)
        else:
            s0 = sg.slab_start(c)
            if cuda:
                k.gemm_planes(op_rows, wpl, trans_b=True, out=buf[s0:s0 + (b_row - a)], **kw)
            else:
                k.gemm(op_rows, w, trans_b=True, out=buf[s0:s0 + (b_row - a)], **kw)

    # ---- layer loop
    pending = None             # (early_done, late_done) events of a replica whose slabs are still landing
    h_rep, h_op = None, None   # replica (gather input) / operand (projection input) of the current h
    if not proj_first(0):
        d0 = layers[0][0].shape[1]
        if feats_local is not None:  # own rows -> gather format, chunk-wise exchange of the replica
            h_rep = new_replica(("xq",), d0)
            for c in range(C):
                a, e = sg.chunk_rows(c)
                s0 = sg.slab_start(c)
                if e > a:
                    if cuda:
                        k.quantize_q24(feats_local[a:e], out=k.Q24(h_rep.data[s0:s0 + e - a], d0))
                    else:
                        h_rep[s0:s0 + e - a] = feats_local[a:e, :d0]
                xch.chunk(replica_2d(h_rep), c)
            mark("features fp32 -> q24 (own rows)")
            xch.wait(f"exchange features ({d0} wide)")
        elif cuda:
            h_rep = new_replica(("xq",), d0)
            k.quantize_q24(feats_pad, out=h_rep)
            mark("features fp32 -> q24")
        else:
            h_rep = feats_pad
    c_out = layers[-1][0].shape[0]
    out = _cached(sg, ("out",), lambda: torch.zeros(sg.total_rows, c_out, dtype=torch.float32, device=dev))

    for l, (w, b) in enumerate(layers):
        last = l == L - 1
        scale, shift = (norms[l] if (norms and not last) else (None, None))
        relu = 0 if last else 1
        d_out, d_in = w.shape
        dpad = (d_out + 3) // 4 * 4
        if proj_first(l):
            wp, bp = _pad_rows(w, b, dpad)
            wpl = k.split_planes(wp) if cuda else None
            z_rep = new_replica(("z", l), dpad)
            if h_op is None:  # first layer: the owned rows of the input features
                h_op = new_operand(("x_op",), d_in)
                mine = (feats_local if feats_local is not None
                        else sg.local_rows_of(feats_pad))[:, :d_in].contiguous()
                if cuda:
                    h_op = k.split_planes(mine)
                else:
                    h_op[: sg.rows] = mine
            two_phase = (cuda and world > 1 and _push_mode() and _push_engine() == "sm"
                         and _two_pass(world, "TWO_PASS_Z"))
            z_events = []
            for c in range(C):
                a, e = sg.chunk_rows(c)
                if e > a:
                    project(rows_of(h_op, a, e), wp, wpl, None, None, None, 0, ("replica", z_rep), a, e, c)
                if two_phase:
                    xch.chunk_two_phase(replica_2d(z_rep), c, z_events)
                else:
                    xch.chunk(replica_2d(z_rep), c)
            mark(f"L{l} gemm {d_in}->{dpad}")
            z_init = None
            if two_phase:
                # first pass over the sources already here while the late slabs of z are in flight
                z_early, z_late = xch.flush_two_phase(replica_2d(z_rep), z_events)
                (pe, ie), (pl_, il) = sg.split_by_owner()
                z_init = _cached(sg, ("zpartial", l), lambda: torch.empty(max(sg.rows, 1), (dpad + 7) // 8 * 8,
                                                                         dtype=torch.float32, device=dev))
                torch.cuda.current_stream().wait_event(z_early)
                k.spmm(pe, ie, z_rep, d=dpad, out=z_init[: sg.rows])
                mark(f"L{l} spmm d={dpad}, pass 1 (own + early source blocks)")
                torch.cuda.current_stream().wait_event(z_late)
                mark(f"L{l} wait for the late source blocks")
                g_ptr, g_idx = pl_, il
            else:
                xch.wait(f"L{l} exchange z ({dpad} wide)")
                g_ptr, g_idx = sg.indptr, sg.indices
            init_rows = (lambda a, e: None) if z_init is None else (lambda a, e: z_init[a:e])
            nxt_pf = (not last) and proj_first(l + 1)
            if last:
                # bias (+ log_softmax) fused into the gather epilogue, straight into the padded output
                for c in range(C):
                    a, e = sg.chunk_rows(c)
                    if e <= a:
                        continue
                    s0 = sg.slab_start(c)
                    if cuda and log_softmax:
                        k.spmm(g_ptr[a:e + 1], g_idx, z_rep, d=dpad, out=out[s0:s0 + (e - a)],
                               dst_scale=sg.inv_deg1[a:e], bias=bp, log_softmax=d_out, acc_init=init_rows(a, e))
                    elif cuda:
                        y = k.spmm(g_ptr[a:e + 1], g_idx, z_rep, d=dpad,
                                   dst_scale=sg.inv_deg1[a:e], bias=bp, acc_init=init_rows(a, e))
                        out[s0:s0 + (e - a)] = y[:, :d_out]
                    else:
                        y = k.spmm_csr(sg.indptr[a:e + 1], sg.indices, z_rep, d=dpad,
                                       dst_scale=sg.inv_deg1[a:e], bias=bp)[:, :d_out]
                        out[s0:s0 + (e - a)] = torch.log_softmax(y, 1) if log_softmax else y
                mark(f"L{l} spmm d={dpad}" + (" +log_softmax" if log_softmax else ""))
                h_rep = h_op = None
            else:
                y_op = new_operand(("y_op", l), dpad)
                if cuda:
                    k.spmm(g_ptr, g_idx, z_rep, d=dpad, out_planes=rows_of(y_op, 0, sg.rows),
                           dst_scale=sg.inv_deg1, bias=bp, col_scale=scale, col_shift=shift, relu=relu,
                           acc_init=None if z_init is None else z_init[: sg.rows])
                else:
                    k.spmm_csr(sg.indptr, sg.indices, z_rep, d=dpad, out=y_op[: sg.rows],
                               dst_scale=sg.inv_deg1, bias=bp, col_scale=scale, col_shift=shift, relu=relu)
                mark(f"L{l} spmm d={dpad}")
                h_op, h_rep = (k.Planes(y_op.hi, y_op.lo, d_out) if cuda else y_op[:, :d_out]), None
                if not nxt_pf:  # an aggregate-first successor needs the replica of this output
                    h_rep = new_replica(("yrep", l), d_out)
                    for c in range(C):
                        a, e = sg.chunk_rows(c)
                        s0 = sg.slab_start(c)
                        if e > a:
                            if cuda:
                                k.quantize_q24(rows_of(h_op, a, e).float(), out=k.Q24(h_rep.data[s0:s0 + e - a], d_out))
                            else:
                                h_rep[s0:s0 + e - a] = h_op[a:e]
                        xch.chunk(replica_2d(h_rep), c)
                    xch.wait(f"L{l} exchange output ({d_out} wide)")
        else:
            wpl = k.split_planes(w) if cuda else None
            agg = new_operand(("agg", l), d_in)
            nxt_pf = (not last) and proj_first(l + 1)
            ldp = (d_in + 7) // 8 * 8
            if cuda and world > 1 and not last and not nxt_pf and _push_mode() and \
                    _replicate_projection() and 4 * ldp < k.Q24.row_bytes(d_out):
                # exchange the narrow aggregated operand, project ALL rows on every rank
                raw, _, _ = _symm_replica(sg, ("aggrep", l), sg.total_rows, 4 * ldp, dev, group)
                both = raw.view(torch.int16)                       # row = [hi (ldp) | lo (ldp)]
                hi, lo = both[:, :ldp], both[:, ldp:]
                for c in range(C):
                    a, e = sg.chunk_rows(c)
                    s0 = sg.slab_start(c)
                    if e > a:
                        k.spmm(sg.indptr[a:e + 1], sg.indices, h_rep, d=d_in,
                               out_planes=k.Planes(hi[s0:s0 + e - a], lo[s0:s0 + e - a], d_in),
                               dst_scale=sg.inv_deg1[a:e])
                    xch.chunk(raw, c)
                mark(f"L{l} spmm d={d_in}")
                xch.wait(f"L{l} exchange aggregated operand ({d_in} wide, planes)")
                y_rep = _cached(sg, ("yrep_local", l), lambda: k.Q24.empty(sg.total_rows, d_out, dev, zero=True))
                k.gemm_planes_q24(k.Planes(hi, lo, d_in), wpl, out=y_rep, bias=b, col_scale=scale,
                                  col_shift=shift, relu=relu)
                mark(f"L{l} gemm {d_in}->{d_out} (all {sg.total_rows} rows, replicated)")
                h_rep, h_op = y_rep, None
                continue
            if last:
                dest = ("final", out)
            elif nxt_pf:
                dest = ("operand", new_operand(("y_op", l), d_out))
            else:
                dest = ("replica", new_replica(("yrep", l), d_out))
            two_phase = (dest[0] == "replica" and cuda and world > 1 and _push_mode()
                         and _push_engine() == "sm" and _two_pass(world))
            pend_events = []
            if pending is not None:
                # this layer's input replica is still arriving: first pass over the sources that are
                # already here (own rows + early owners), for ALL local rows, into fp32 partial sums
                (pe, ie), (pl_, il) = sg.split_by_owner()
                partial = _cached(sg, ("partial", l), lambda: torch.empty(max(sg.rows, 1), (d_in + 7) // 8 * 8,
                                                                         dtype=torch.float32, device=dev))
                torch.cuda.current_stream().wait_event(pending[0])
                k.spmm(pe, ie, h_rep, d=d_in, out=partial[: sg.rows])
                mark(f"L{l} spmm d={d_in}, pass 1 (own + early source blocks)")
                torch.cuda.current_stream().wait_event(pending[1])
                mark(f"L{l} wait for the late source blocks")
            for c in range(C):
                a, e = sg.chunk_rows(c)
                if e > a:
                    if pending is not None:   # second pass: the late blocks, on top of the partial sums
                        k.spmm(pl_[a:e + 1], il, h_rep, d=d_in, out_planes=rows_of(agg, a, e),
                               dst_scale=sg.inv_deg1[a:e], acc_init=partial[a:e])
                    else:
                        aggregate(h_rep, d_in, a, e, agg)
                    project(rows_of(agg, a, e), w, wpl, b, scale, shift, relu, dest, a, e, c)
                if dest[0] == "replica":
                    if two_phase:
                        xch.chunk_two_phase(replica_2d(dest[1]), c, pend_events)
                    else:
                        xch.chunk(replica_2d(dest[1]), c)
            mark(f"L{l} spmm d={d_in}" + (", pass 2" if pending is not None else "") + f" + gemm {d_in}->{d_out}")
            pending = None
            if dest[0] == "replica" and two_phase and (l + 1 < L) and not proj_first(l + 1):
                pending = xch.flush_two_phase(replica_2d(dest[1]), pend_events)
                h_rep, h_op = dest[1], None
            elif dest[0] == "replica" and two_phase:
                done = xch.flush_two_phase(replica_2d(dest[1]), pend_events)
                torch.cuda.current_stream().wait_event(done[1])
                mark(f"L{l} exchange output ({d_out} wide)")
                h_rep, h_op = dest[1], None
            elif dest[0] == "replica":
                xch.wait(f"L{l} exchange output ({d_out} wide)")
                h_rep, h_op = dest[1], None
            elif dest[0] == "operand":
                h_rep, h_op = None, dest[1]
            else:
                if log_softmax:
                    for c in range(C):
                        a, e = sg.chunk_rows(c)
                        s0 = sg.slab_start(c)
                        if e > a:
                            k.log_softmax(out[s0:s0 + e - a], out=out[s0:s0 + e - a])
                    mark("log_softmax")
    if gather_output:  # otherwise every rank keeps only its own rows valid
        for c in range(C):
            xch.chunk(out, c)
        xch.wait("exchange output")
    return out
