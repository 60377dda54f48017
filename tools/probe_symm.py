"""Probe: torch symmetric memory on this box (peer pointers, signal pads, multicast), 2+ ranks."""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
t.fill_(float(rank))
hdl.barrier(channel=0)
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1 << 20,), torch.float32)
pt[:16] = 100.0 + rank           # P2P store into the neighbour
torch.cuda.synchronize()
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(f"rank {rank}: ptrs={[hex(p) for p in hdl.buffer_ptrs]} sig={[hex(p) for p in hdl.signal_pad_ptrs][:2]} "
      f"multicast={hdl.has_multicast_support and hex(hdl.multicast_ptr)} sigsize={hdl.signal_pad_size} "
      f"local[:2]={t[:2].tolist()} local[16:18]={t[16:18].tolist()}", flush=True)
# bandwidth of a peer copy through the symmetric mapping
big = symm.empty(256 << 20, dtype=torch.uint8, device=dev)
h2 = symm.rendezvous(big, dist.group.WORLD)
src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dstp = h2.get_buffer(peer, (256 << 20,), torch.uint8)
for _ in range(2):
    dstp.copy_(src)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dstp.copy_(src)
e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: peer copy {5 * 256 / 1024 / (e0.elapsed_time(e1) / 1e3):.1f} GiB/s", flush=True)
dist.barrier(); dist.destroy_process_group()
