"""Multi-GPU (needs >= 2 visible GPUs; skipped otherwise): the dst-row sharded SAGE forward with one
NCCL all-gather per layer equals the single-GPU forward on the same inputs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world, device_id=dev)
    from glnn_b200 import dist_teacher as DT, graph as G, ops
    from glnn_b200.models import Model
    from glnn_b200.workloads import randomise_bn_, synthetic_graph
    n = 50000
    g = synthetic_graph(n, 600000, mirror=True, self_loops=False, device=dev, seed=0)
    torch.manual_seed(0)
    model = randomise_bn_(Model(dict(model_name="SAGE", num_layers=3, feat_dim=100, hidden_dim=256,
                                     label_dim=47, dropout_ratio=0.5, norm_type="batch",
                                     device=dev))).eval()
    feats = torch.randn(n, 100, generator=torch.Generator().manual_seed(1)).to(dev)
    sg = DT.ShardedGraph(g, rank, world)
    enc = model.encoder
    layers = [(c.fc_neigh.weight.detach(), c.fc_neigh.bias.detach()) for c in enc.layers]
    norms = [ops.bn_fold(b.weight, b.bias, b.running_mean, b.running_var, b.eps) for b in enc.norms]
    with torch.no_grad():
        out = sg.from_padded(DT.sage_forward_sharded(sg, sg.to_padded(feats), layers, norms))
        ref = enc.inference(G.FullNeighborLoader(g), feats, log_softmax=True)
    err = float((out - ref).abs().max() / ref.abs().max())
    q.put((rank, err))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_forward_matches_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err < 1e-5, (rank, err)
