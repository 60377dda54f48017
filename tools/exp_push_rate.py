"""What does the all-to-all slab exchange of the sharded teacher reach on this box?  Every rank pushes
`mb` MB to every peer's symmetric buffer simultaneously (the pattern of the layer exchange), with the
SM-driven kernel (glnn_peer_push, several CTA counts) and with the copy engines, alone and while an
HBM-saturating copy loop runs on another stream (standing in for the neighbour gather).
Run:  torchrun --nproc-per-node N tools/exp_push_rate.py   -> one JSON line per variant (rank 0)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

from glnn_b200 import ops

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
mb = int(os.environ.get("PUSH_MB", "200"))
n = mb * 1000 * 1000
buf = symm.empty(world * n, dtype=torch.uint8, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
buf.zero_()
peers = {r: hdl.get_buffer(r, (world * n,), torch.uint8) for r in range(world) if r != rank}
mine = buf[rank * n:(rank + 1) * n]
order = [(rank + i) % world for i in range(1, world)]
side = torch.cuda.Stream()
streams = [torch.cuda.Stream(priority=-1) for _ in range(world - 1)]
big_a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
big_b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)


def push_sm(ctas):
    ops.peer_push(mine, [peers[p][rank * n:(rank + 1) * n] for p in order], ctas)


def push_ce():
    ev = torch.cuda.Event()
    ev.record()
    for i, p in enumerate(order):
        streams[i].wait_event(ev)
        with torch.cuda.stream(streams[i]):
            peers[p][rank * n:(rank + 1) * n].copy_(mine, non_blocking=True)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)


def timed(fn, load, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    if load:
        with torch.cuda.stream(side):
            for _ in range(40):
                big_b.copy_(big_a)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out_gb = (world - 1) * n / 1e9
for load in (False, True):
    for name, fn in [("sm 16 CTAs", lambda: push_sm(16)), ("sm 32 CTAs", lambda: push_sm(32)),
                     ("sm 64 CTAs", lambda: push_sm(64)), ("sm 148 CTAs", lambda: push_sm(148)),
                     ("copy engines", push_ce)]:
        ms = timed(fn, load)
        if rank == 0:
            print(json.dumps(dict(world=world, mb_per_peer=mb, variant=name, hbm_load=load, ms=round(ms, 3),
                                  out_GB_per_rank=round(out_gb, 3),
                                  TBps_out_per_rank=round(out_gb / ms, 3))), flush=True)
dist.barrier()
dist.destroy_process_group()
