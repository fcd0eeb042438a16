"""One dense layer through the C ABI, for ncu.  Usage: python tools/profile_linear.py m n k [bf16|f32] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mil_b200
from mil_b200 import functional as F
m, n, k = (int(v) for v in sys.argv[1:4])
dtype = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else torch.bfloat16
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
x = torch.randn(m, k, device="cuda").to(dtype)
W = (torch.randn(n, k, device="cuda") / k ** 0.5).to(dtype)
b = torch.randn(n, device="cuda")
for _ in range(reps):
    y = F.linear(x, W, b)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), mil_b200.launch_count())
if os.environ.get("TRACE"):
    from mil_b200 import _lib as L
    tr = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.check(L.lib().milb200_debug_trace(L.ptr(tr)), "trace")
    for _ in range(3):
        tr.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = F.linear(x, W, b); e1.record()
        torch.cuda.synchronize()
        t = tr.cpu().tolist()
        print("event us", e0.elapsed_time(e1) * 1e3, "cycles since entry:", [(i, t[i] - t[0]) for i in range(1, 10)])
    L.lib().milb200_debug_trace(None)
