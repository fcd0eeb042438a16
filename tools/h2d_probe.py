"""H2D bandwidth probe: pinned host -> device, one stream vs two concurrent streams, several chunk sizes."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(nstreams, chunks):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    per = n // chunks
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(5):
        for c in range(chunks):
            with torch.cuda.stream(streams[c % nstreams]):
                d[c * per:(c + 1) * per].copy_(h[c * per:(c + 1) * per], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"streams={nstreams} chunks={chunks}: {n / dt / 1e9:.1f} GB/s", flush=True)
run(1, 1); run(1, 1)
run(1, 8)
run(2, 2); run(2, 8); run(4, 16)
