"""Smallest program that runs the fused aggregator step (BASELINE configs[2] shape) — the thing ncu wraps.
Usage: python tools/profile_fusion.py [N] [T] [dtype: f32|bf16] [iters]"""
import os
import sys
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mil_b200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dtype = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float32
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
args = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                 aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
torch.manual_seed(1234)
m = mil_b200.get_model(args).cuda().to(dtype).train(False)
x_ct = torch.randn(1, 512, 160, 1, 1, device="cuda", dtype=dtype)
x_p = torch.randn(1, N, 768, device="cuda", dtype=dtype)
x_t = (torch.randn(1, T, 512, device="cuda") * 0.05).to(dtype)
label = torch.tensor([[0.0, 1.0]], device="cuda")
for it in range(iters):
    torch.cuda.synchronize()
    l0 = mil_b200.launch_count()
    m.zero_grad(set_to_none=True)
    prob, a, b = m([x_ct, x_p], x_t)
    loss = torch.nn.functional.binary_cross_entropy(prob.float(), label) + \
        mil_b200.clip_loss.cosine_embedding_loss(a.squeeze(0), b.squeeze(0)).float()
    loss.backward()
    torch.cuda.synchronize()
    print("iter", it, "launches", mil_b200.launch_count() - l0, "loss", float(loss))
