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
def step():
    m.zero_grad(set_to_none=True)
    p, a, b = m([x_ct, x_p], x_t)
    (p.sum() + (a * b).sum()).backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(30): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue()[:9000])
