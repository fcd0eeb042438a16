"""Attention kernels through the C ABI only (no autograd, no allocations in the loop): warm kernel times at the fusion
path's shapes.  bytes = Q + K + V + O (+ dO, dQ, dK, dV in backward) once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mil_b200
from mil_b200 import _lib as L

lib = L.lib()
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

H, c = 8, 32
for dtype in (torch.bfloat16, torch.float32):
    code = L.dtype_code(torch.empty(0, dtype=dtype))
    for nq, nk in ((1, 15592), (10, 15592), (15592, 1), (15592, 10), (10, 160), (160, 10)):
        mk = lambda n: torch.randn(n, H * c, device="cuda").to(dtype)
        q, k, v, do = mk(nq), mk(nk), mk(nk), mk(nq)
        o, dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        lse = torch.empty(H * nq, dtype=torch.float32, device="cuda")
        wf = torch.empty(max(lib.milb200_attention_workspace_bytes(nq, nk, H, c, 0), 256), dtype=torch.uint8, device="cuda")
        wb = torch.empty(max(lib.milb200_attention_workspace_bytes(nq, nk, H, c, 1), 256), dtype=torch.uint8, device="cuda")
        st = L.stream_ptr()
        f = lambda: lib.milb200_attention_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(lse), nq, nk, H, c, code, L.ptr(wf), wf.numel(), st)
        assert f() == 0
        bw = lambda: lib.milb200_attention_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(o), L.ptr(lse), L.ptr(do), L.ptr(dq), L.ptr(dk), L.ptr(dv), nq, nk, H, c, code, L.ptr(wb), wb.numel(), st)
        assert bw() == 0
        tf, tb = timeit(f), timeit(bw)
        e = q.element_size()
        bf = (2 * nq + 2 * nk) * H * c * e
        bb = (3 * nq + 4 * nk + nq) * H * c * e
        print(f"{str(dtype)[6:]:9s} nq={nq:6d} nk={nk:6d}: fwd {tf:6.1f} us ({bf / tf / 1e3:6.0f} GB/s)   bwd {tb:6.1f} us ({bb / tb / 1e3:6.0f} GB/s)", flush=True)
