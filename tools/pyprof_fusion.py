import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mil_b200
from argparse import Namespace
ARGS = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                 aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
dtype = torch.bfloat16
m = mil_b200.get_model(ARGS).cuda().to(dtype).train(False)
x_ct = torch.randn(1, 512, 160, 1, 1, device="cuda", dtype=dtype)
x_p = torch.randn(1, 15592, 768, device="cuda", dtype=dtype)
x_t = (torch.randn(1, 1, 512, device="cuda") * 0.05).to(dtype)
plist = list(m.parameters())
def step():
    for p_ in plist: p_.grad = None
    p, a, b = m([x_ct, x_p], x_t)
    (p.sum() + (a * b).sum()).backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(30): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue()[:6000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(40); print(s.getvalue()[:8000])

# host enqueue time vs device time for the same loop
import time
for N in (1000, 15592):
    x_p = torch.randn(1, N, 768, device="cuda", dtype=dtype)
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): step()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"N={N}: host enqueue {1e3*(t1-t0)/50:.3f} ms/step, wall {1e3*(t2-t0)/50:.3f} ms/step, device span {e0.elapsed_time(e1)/50:.3f} ms/step")
