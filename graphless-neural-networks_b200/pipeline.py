"""Host-facing throughput path for the teacher forward: host (pinned) CSR + features in, host
log-probabilities out, with the PCIe copies of neighbouring steps overlapped with compute.

A step needs all of its inputs before the first aggregation (random row gathers), so nothing inside
ONE forward can overlap with its own upload; across steps, however, step i+1's upload and step i-1's
download run on copy streams while step i computes (double-buffered device inputs, full-duplex
PCIe).  Every step still copies its own inputs and reads back its own result."""
import torch

from .graph import CSRGraph, FullNeighborLoader


class HostTeacherPipeline:
    def __init__(self, encoder, n, indptr_like, indices_like, feat_dim, label_dim, device, depth=2):
        self.enc, self.n, self.dev, self.depth = encoder, n, device, depth
        self.h2d, self.d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
        self.slots = []
        for _ in range(depth):
            slot = dict(
                indptr=torch.empty_like(indptr_like, device=device),
                indices=torch.empty_like(indices_like, device=device),
                feats=torch.empty(n, feat_dim, dtype=torch.float32, device=device),
                uploaded=torch.cuda.Event(), computed=torch.cuda.Event(), downloaded=torch.cuda.Event())
            slot["loader"] = FullNeighborLoader(CSRGraph(slot["indptr"], slot["indices"], n))
            self.slots.append(slot)
        self.step = 0
        self.label_dim = label_dim

    def submit(self, h_indptr, h_indices, h_feats, h_out):
        """Enqueues one forward: upload -> SAGE.inference + log_softmax -> download into h_out
        (pinned).  Returns immediately; call drain() before reading h_out."""
        s = self.slots[self.step % self.depth]
        cur = torch.cuda.current_stream(self.dev)
        if self.step >= self.depth:
            self.h2d.wait_event(s["computed"])  # the slot's previous forward has consumed its inputs
        with torch.cuda.stream(self.h2d):
            s["indptr"].copy_(h_indptr, non_blocking=True)
            s["indices"].copy_(h_indices, non_blocking=True)
            s["feats"].copy_(h_feats, non_blocking=True)
            s["uploaded"].record(self.h2d)
        cur.wait_event(s["uploaded"])
        with torch.no_grad():
            out = self.enc.inference(s["loader"], s["feats"], log_softmax=True)
        s["computed"].record(cur)
        self.d2h.wait_event(s["computed"])
        with torch.cuda.stream(self.d2h):
            out.record_stream(self.d2h)
            h_out.copy_(out, non_blocking=True)
            s["downloaded"].record(self.d2h)
        self.step += 1

    def drain(self):
        cur = torch.cuda.current_stream(self.dev)
        for s in self.slots:
            cur.wait_event(s["downloaded"])
            cur.wait_event(s["computed"])


class HostShardedTeacherPipeline:
    """The same host-facing path for the dst-row sharded forward (one instance per rank): every step
    uploads THIS rank's CSR slice and THIS rank's feature rows (1/G of the features -- the input
    replica is assembled over NVLink by sage_forward_sharded(feats_local=...)), runs the sharded
    forward and downloads this rank's rows of the log-probabilities.  Uploads / downloads of
    neighbouring steps overlap with compute exactly as in HostTeacherPipeline."""

    def __init__(self, sg, layers, norms, feat_dim, label_dim, device, group=None, depth=2):
        from . import dist_teacher as DT
        self.sg, self.layers, self.norms, self.group = sg, layers, norms, group
        self.dev, self.depth, self.step = device, depth, 0
        self.h2d, self.d2h = torch.cuda.Stream(device), torch.cuda.Stream(device)
        # with the two-pass exchange the aggregations read the CSR split by source owner: the host ships
        # THAT form instead of the merged CSR (same edges, two index arrays, each edge once), so every
        # gather of a step runs on data uploaded in that step
        self.two_pass = bool(device.type == "cuda" and sg.world > 1 and DT._two_pass(sg.world))
        self.slots = []
        for _ in range(depth):
            slot = dict(
                indptr=torch.empty_like(sg.indptr), indices=torch.empty_like(sg.indices),
                feats=torch.empty(sg.rows, feat_dim, dtype=torch.float32, device=device),
                out=torch.empty(sg.rows, label_dim, dtype=torch.float32, device=device),
                uploaded=torch.cuda.Event(), computed=torch.cuda.Event(), downloaded=torch.cuda.Event())
            if self.two_pass:
                slot["split"] = tuple(torch.empty_like(t) for pair in sg.split_by_owner() for t in pair)
            self.slots.append(slot)

    def host_split(self):
        """Pinned host copies of the CSR split by source owner, (early indptr, early indices, late indptr,
        late indices), or None when the exchange is consumed in one pass."""
        if not self.two_pass:
            return None
        return tuple(t.cpu().pin_memory() for pair in self.sg.split_by_owner() for t in pair)

    def submit(self, h_indptr, h_indices, h_feats, h_out, h_split=None):
        """h_*: pinned host tensors (this rank's relabelled CSR slice, its feature rows [rows, F]);
        h_out: pinned [rows, C]; h_split: host_split() when the two-pass exchange is active (then
        h_indptr / h_indices are not read).  Returns immediately; drain() before reading h_out."""
        if self.two_pass and h_split is None:
            raise ValueError("two-pass exchange: pass h_split=pipe.host_split()")
        from . import dist_teacher as DT
        sg, s = self.sg, self.slots[self.step % self.depth]
        cur = torch.cuda.current_stream(self.dev)
        if self.step >= self.depth:
            self.h2d.wait_event(s["computed"])   # the slot's previous forward has consumed its inputs
            cur.wait_event(s["downloaded"])      # ... and its output rows have left the device
        with torch.cuda.stream(self.h2d):
            if self.two_pass:   # every edge travels once: the owner-split form replaces the merged CSR
                for dst, src in zip(s["split"], h_split):
                    dst.copy_(src, non_blocking=True)
            else:
                s["indptr"].copy_(h_indptr, non_blocking=True)
                s["indices"].copy_(h_indices, non_blocking=True)
            s["feats"].copy_(h_feats, non_blocking=True)
            s["uploaded"].record(self.h2d)
        cur.wait_event(s["uploaded"])
        keep = (sg.indptr, sg.indices, sg.__dict__.get("_split"))
        sg.indptr, sg.indices = s["indptr"], s["indices"]
        if self.two_pass:
            sp = s["split"]
            sg._split = ((sp[0], sp[1]), (sp[2], sp[3]))
        try:
            with torch.no_grad():
                out = DT.sage_forward_sharded(sg, None, self.layers, self.norms, group=self.group,
                                              gather_output=False, feats_local=s["feats"],
                                              split_only=self.two_pass)
        finally:
            sg.indptr, sg.indices = keep[0], keep[1]
            if self.two_pass:
                sg._split = keep[2]
        for c in range(sg.chunks):   # this rank's rows of the padded output, back in local order
            a, e = sg.chunk_rows(c)
            if e > a:
                s0 = sg.slab_start(c)
                s["out"][a:e].copy_(out[s0:s0 + e - a], non_blocking=True)
        s["computed"].record(cur)
        self.d2h.wait_event(s["computed"])
        with torch.cuda.stream(self.d2h):
            h_out.copy_(s["out"], non_blocking=True)
            s["downloaded"].record(self.d2h)
        self.step += 1

    def drain(self):
        cur = torch.cuda.current_stream(self.dev)
        for s in self.slots:
            cur.wait_event(s["downloaded"])
            cur.wait_event(s["computed"])
