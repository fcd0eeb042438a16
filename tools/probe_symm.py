"""Probe (2+ GPUs, torchrun): what peer-memory plumbing works on this box — torch symmetric memory (multicast?), CUDA IPC."""
import os
import sys
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    dist.barrier()
    torch.cuda.synchronize()
    peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    print(f"[rank {rank}] symm ok: world {h.world_size} multicast_ptr {getattr(h, 'multicast_ptr', None)} "
          f"peer value {float(peer[0])} buffer_ptrs {len(h.buffer_ptrs)} signal_pads {len(h.signal_pad_ptrs)}", flush=True)
except Exception as e:
    print(f"[rank {rank}] symm FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
try:
    x = torch.full((1024,), float(rank + 1), device=dev)
    torch.cuda.synchronize()
    can = [torch.cuda.can_device_access_peer(dev.index, j) for j in range(torch.cuda.device_count()) if j != dev.index]
    print(f"[rank {rank}] can_device_access_peer: {can}", flush=True)
except Exception as e:
    print(f"[rank {rank}] peer probe failed: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
