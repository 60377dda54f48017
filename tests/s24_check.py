"""Standalone checker of the EXPERIMENTAL sparse-row (s24) kernels, run by
tests/test_zz_next_rows_gpu.py in a SUBPROCESS (a device fault in an unverified kernel must not
poison the pytest process):  python tests/s24_check.py <d> <zero_frac>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import relerr  # noqa: E402


def main(d, zero_frac):
    from glnn_b200 import ops
    from glnn_b200.workloads import synthetic_graph
    dev = torch.device("cuda:0")
    n = 20000
    g = synthetic_graph(n, 300000, mirror=True, self_loops=False, device=dev, seed=3)  # hubs > 1024 edges
    gen = torch.Generator().manual_seed(d)
    x = torch.randn(n, d, generator=gen)
    if zero_frac > 0:
        x = torch.relu(x - float(torch.quantile(x.flatten()[:200000], zero_frac)))
    xq = ops.quantize_q24(x.to(dev))
    s = ops.compact_s24(xq)
    # ---- layout
    vals = xq.float().cpu().numpy()
    words = s.data.cpu().numpy().view(np.uint32)
    nnz = (vals != 0).sum(1)
    assert int(s.cap.item()) == int(nnz.max()), (int(s.cap.item()), int(nnz.max()))
    assert words.shape[1] % 32 == 0 and words.shape[1] >= d
    for r in (0, 1, n // 2, n - 1, int(nnz.argmax())):
        cols = np.flatnonzero(vals[r])
        want = vals[r, cols].astype(np.float32).view(np.uint32) | cols.astype(np.uint32)
        assert np.array_equal(words[r, :cols.size], want), r
        assert not words[r, cols.size:].any(), r
    # ---- aggregation
    bias = torch.randn(d, generator=gen).to(dev)
    scale, shift = (torch.rand(d, generator=gen) + 0.5).to(dev), torch.randn(d, generator=gen).to(dev)
    kw = dict(self_add=True, mean_plus_one=True, bias=bias, col_scale=scale, col_shift=shift, relu=1)
    want_p = ops.spmm(g.indptr, g.indices, xq, out_planes=ops.new_planes(n, d, dev), **kw)
    got_p = ops.spmm(g.indptr, g.indices, xq, out_planes=ops.new_planes(n, d, dev), s24=s, **kw)
    want, got = want_p.float().cpu(), got_p.float().cpu()
    assert relerr(got, want) < 1e-6, relerr(got, want)
    small = (g.in_degrees().cpu() <= 1024)
    assert small.sum() < n and torch.equal(got[small], want[small])
    want_f = ops.spmm(g.indptr, g.indices, xq, **kw)
    got_f = ops.spmm(g.indptr, g.indices, xq, s24=s, **kw)
    assert relerr(got_f.cpu(), want_f.cpu()) < 1e-6
    torch.cuda.synchronize()
    print("s24 ok", d, zero_frac, "cap", int(s.cap.item()))


if __name__ == "__main__":
    main(int(sys.argv[1]), float(sys.argv[2]))
