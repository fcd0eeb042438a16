"""Multi-GPU H2D probe (torchrun, one rank per GPU): what caps the end-to-end rate at N = 4 / 8?  Every rank copies 1 GiB
pinned host -> its device, (a) alone, one rank at a time, (b) all ranks at once, (c) all at once with each rank pinned to
its own slice of the host cores / the GPU's NUMA node before the pinned allocation.  Prints GB/s per rank."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 30


def measure(h, d, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


def gather(x):
    t = torch.tensor([x], device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [round(float(o), 1) for o in out]


d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
measure(h, d, 1)
alone = 0.0
for r in range(world):                      # (a) one rank at a time
    dist.barrier()
    if r == rank:
        alone = measure(h, d)
dist.barrier()
a = gather(alone)
dist.barrier()
together = measure(h, d)                    # (b) all ranks at once
b = gather(together)
cores = sorted(os.sched_getaffinity(0))
per = max(1, len(cores) // world)
os.sched_setaffinity(0, set(cores[rank * per:(rank + 1) * per]))
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()      # first touch from the pinned cores
h2.fill_(1)
measure(h2, d, 1)
dist.barrier()
pinned = measure(h2, d)                     # (c) all at once, core-pinned allocation
c = gather(pinned)
if rank == 0:
    print(f"host cores visible: {len(cores)}; ranks: {world}")
    print(f"(a) one rank at a time      GB/s per rank: {a}")
    print(f"(b) all ranks concurrently  GB/s per rank: {b}  sum {sum(b):.1f}")
    print(f"(c) concurrent, core-pinned GB/s per rank: {c}  sum {sum(c):.1f}")
    try:
        print(open("/sys/devices/system/node/online").read().strip(), "NUMA nodes online")
    except Exception:
        pass
dist.destroy_process_group()
