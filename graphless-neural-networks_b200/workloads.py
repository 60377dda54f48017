"""Synthetic inputs with the shapes of the reference's datasets: there is no
network for the real ogbn-* / CPF files, so benchmarks and full-size tests use graphs with the same
node / edge counts and a skewed in-degree distribution, random features and random-init weights."""
import torch

from .graph import CSRGraph

SHAPES = {
    # name: (nodes, directed edges before mirroring, feat_dim, classes, self_loops, split sizes)
    "ogbn-arxiv": dict(n=169343, e_raw=1166243, feat=128, classes=40, self_loops=True,
                       split=(90941, 29799, 48603), hidden=256, batch_size=512),
    "ogbn-products": dict(n=2449029, e_raw=61859140, feat=100, classes=47, self_loops=False,
                          split=(196615, 39323, 2213091), hidden=256, batch_size=4096),
    "cora": dict(n=2485, e_raw=5069, feat=1433, classes=7, self_loops=True, split=(140, 210, 2135),
                 hidden=64, batch_size=512),
}


def synthetic_edges(n, e_raw, mirror=True, self_loops=False, device="cpu", seed=0):
    """src uniform, dst = floor(n * u^2) (hubs at low ids); optionally mirrored without dedup, with
    self-loops removed and one self-loop per node added (dataloader.py:74-77 for ogbn-arxiv)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, n, (e_raw,), device=device, generator=gen)
    u = torch.rand(e_raw, device=device, generator=gen, dtype=torch.float64)
    dst = (u * u * n).floor().to(torch.int64).clamp_(max=n - 1)
    del u
    if mirror:
        src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    if self_loops:
        keep = src != dst
        loops = torch.arange(n, device=device)
        src, dst = torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])
    return src, dst


def synthetic_graph(n, e_raw, mirror=True, self_loops=False, device="cpu", seed=0):
    src, dst = synthetic_edges(n, e_raw, mirror, self_loops, device, seed)
    return CSRGraph.from_edges(src, dst, n)


def dataset_graph(name, device="cpu", seed=0):
    s = SHAPES[name]
    return synthetic_graph(s["n"], s["e_raw"], True, s["self_loops"], device, seed)


def randomise_bn_(model, seed=0):
    """Non-trivial BatchNorm statistics so that the eval-mode affine is not the identity."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                k = m.num_features
                m.running_mean.copy_(torch.randn(k, generator=gen))
                m.running_var.copy_(torch.rand(k, generator=gen) * 1.5 + 0.5)
                m.weight.copy_(torch.rand(k, generator=gen) + 0.5)
                m.bias.copy_(torch.randn(k, generator=gen))
    return model


def sage_bytes_per_forward(n, e, dims):
    """Algorithmic (compulsory) HBM bytes of one fused SAGE forward, BASELINE.md section 3:
    per layer 4(N+1) + 4E + 4 N d_in + 4 N d_out + 4 d_in d_out + 4 d_out (+ 8 d_out BN affine)."""
    total = 0
    for l in range(len(dims) - 1):
        d_in, d_out = dims[l], dims[l + 1]
        total += 4 * (n + 1) + 4 * e + 4 * n * d_in + 4 * n * d_out + 4 * d_in * d_out + 4 * d_out
        if l != len(dims) - 2:
            total += 8 * d_out
    return total


def sage_gather_bytes(n, e, dims):
    """Secondary diagnostic: bytes moved by the neighbour gathers if nothing hits in L2
    (4 * E * d_agg per layer, d_agg = the narrower side after 4-padding the output)."""
    total = 0
    for l in range(len(dims) - 1):
        d_in, d_out = dims[l], (dims[l + 1] + 3) // 4 * 4
        total += 4 * e * (d_out if d_out < d_in else d_in)
    return total
