import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, mil_b200
from argparse import Namespace
ARGS = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                 aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
for dtype in (torch.bfloat16, torch.float32):
    m = mil_b200.get_model(ARGS).cuda().to(dtype).train(False)
    for N in (1000, 15592):
        x_ct = torch.randn(1, 512, 160, 1, 1, device="cuda", dtype=dtype)
        x_p = torch.randn(1, N, 768, device="cuda", dtype=dtype)
        x_t = (torch.randn(1, 1, 512, device="cuda") * 0.05).to(dtype)
        with torch.no_grad():
            for _ in range(3): m([x_ct, x_p], x_t)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10): m([x_ct, x_p], x_t)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
        print(dtype, N, "fwd: cpu enqueue ms/iter", (t1 - t0) * 100, "total ms/iter", (t2 - t0) * 100, flush=True)
        for _ in range(3):
            m.zero_grad(set_to_none=True); p, a, b = m([x_ct, x_p], x_t); (p.sum() + (a * b).sum()).backward()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            m.zero_grad(set_to_none=True); p, a, b = m([x_ct, x_p], x_t); (p.sum() + (a * b).sum()).backward()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(dtype, N, "fwd+bwd: cpu enqueue ms/iter", (t1 - t0) * 100, "total ms/iter", (t2 - t0) * 100, flush=True)
