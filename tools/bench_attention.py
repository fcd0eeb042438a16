"""Kernel timings of the attention core at the fusion path's shapes (CUDA events, 20 reps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mil_b200
from mil_b200 import functional as F

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

for dtype in (torch.bfloat16, torch.float32):
    for (nq, nk, c) in ((1, 15592, 32), (10, 15592, 32), (15592, 1, 32), (15592, 10, 32), (10, 10, 64), (1, 160, 32), (160, 1, 32)):
        H = 8
        q = torch.randn(nq, H * c, device="cuda").to(dtype).requires_grad_(True)
        k = torch.randn(nk, H * c, device="cuda").to(dtype).requires_grad_(True)
        v = torch.randn(nk, H * c, device="cuda").to(dtype).requires_grad_(True)
        do = torch.randn(nq, H * c, device="cuda").to(dtype)
        with torch.no_grad():
            tf = t(lambda: F.attention_core(q, k, v, H))
        o = F.attention_core(q, k, v, H)
        tb = t(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
        bytes_f = (nq + 2 * nk + nq) * H * c * q.element_size()
        print(f"attention {str(dtype)[6:]} nq={nq} nk={nk} c={c}: fwd {tf:.1f} us ({bytes_f / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us", flush=True)
