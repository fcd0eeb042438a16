"""Host-side cost of the C-ABI entry points (us per call, enqueue only), to tell CPU-bound from GPU-bound paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mil_b200
from mil_b200 import _lib as L, functional as F

def loop(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn()
        if i % 50 == 49: torch.cuda.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

lib = L.lib(); st = L.stream_ptr()
for dtype in (torch.bfloat16, torch.float32):
    code = 1 if dtype == torch.bfloat16 else 0
    for (m, n, k) in ((1, 512, 512), (160, 256, 512), (15592, 256, 512)):
        x = torch.randn(m, k, device="cuda").to(dtype); W = torch.randn(n, k, device="cuda").to(dtype)
        b = torch.randn(n, device="cuda"); y = torch.empty(m, n, device="cuda", dtype=dtype)
        ws = torch.empty(lib.milb200_linear_workspace_bytes(m, n, k, code, 1), dtype=torch.uint8, device="cuda")
        px, pw, pb, py, pws = (L.ptr(t) for t in (x, W, b, y, ws))
        us = loop(lambda: lib.milb200_linear_fwd(px, None, pw, pb, py, m, n, k, 0, code, pws, ws.numel(), st))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(100): lib.milb200_linear_fwd(px, None, pw, pb, py, m, n, k, 0, code, pws, ws.numel(), st)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"linear_fwd {dtype} m={m} n={n} k={k}: {us:.1f} us/call (with periodic sync); burst of 100: enqueue {(t1-t0)*1e4:.1f} us/call, drained {(t2-t0)*1e4:.1f} us/call", flush=True)
    x = torch.randn(1000, 512, device="cuda").to(dtype); g = torch.ones(512, device="cuda"); be = torch.zeros(512, device="cuda")
    y = torch.empty_like(x); mean = torch.empty(1000, device="cuda"); rstd = torch.empty(1000, device="cuda")
    px, pg, pbe, py, pm, pr = (L.ptr(t) for t in (x, g, be, y, mean, rstd))
    print(f"layernorm_fwd {dtype}: {loop(lambda: lib.milb200_layernorm_fwd(px, None, pg, pbe, py, pm, pr, 1000, 512, code, 0, st)):.1f} us/call", flush=True)
    a = torch.randn(1000, 512, device="cuda").to(dtype)
    pa = L.ptr(a)
    print(f"add {dtype}: {loop(lambda: lib.milb200_add(px, pa, py, 1000 * 512, code, st)):.1f} us/call", flush=True)
e = torch.empty(1024, device="cuda")
print(f"torch.empty: {loop(lambda: torch.empty(1024, device='cuda')):.1f} us; python no-op lambda: {loop(lambda: None):.2f} us", flush=True)
