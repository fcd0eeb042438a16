"""Where the HOST time of one fusion step goes (no profiler: perf_counter around segments, enqueue only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mil_b200
from argparse import Namespace
ARGS = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                 aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
dtype = torch.bfloat16
m = mil_b200.get_model(ARGS).cuda().to(dtype).train(False)
plist = list(m.parameters())
x_ct = torch.randn(1, 512, 160, 1, 1, device="cuda", dtype=dtype)
x_p = torch.randn(1, 15592, 768, device="cuda", dtype=dtype)
x_t = (torch.randn(1, 1, 512, device="cuda") * 0.05).to(dtype)
label = torch.tensor([[0.0, 1.0]], device="cuda")
acc = {}
def seg(name, t0):
    t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
def step(rec):
    t = time.perf_counter()
    for p in plist: p.grad = None
    if rec: t = seg("zero_grad", t)
    prob, a, b = m([x_ct, x_p], x_t)
    if rec: t = seg("forward", t)
    loss = torch.nn.functional.binary_cross_entropy(prob.float(), label) + \
        mil_b200.clip_loss.cosine_embedding_loss(a.squeeze(0), b.squeeze(0)).float()
    if rec: t = seg("loss", t)
    loss.backward()
    if rec: t = seg("backward", t)
for _ in range(10): step(False)
torch.cuda.synchronize()
n = 100
t0 = time.perf_counter()
for _ in range(n): step(True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print({k: round(1e3 * v / n, 3) for k, v in acc.items()}, "host ms/step", round(1e3 * (t1 - t0) / n, 3), "wall", round(1e3 * (t2 - t0) / n, 3))
# inside forward: the tape call alone
from mil_b200 import functional as F
tape = m._fusion_tape(True)
xin = [F.ct_tokens(x_ct)[0], m._pe(160, x_t)[0], x_p[0], m._pe(15592, x_t)[0], x_t[0]]
rows = {"T": 1, "Nc": 160, "Np": 15592}
with torch.no_grad():
    for _ in range(5): tape.run(rows, xin)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): tape.run(rows, xin)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("tape.run forward only (no grad): host", round(1e3 * (t1 - t0) / n, 3), "ms, wall", round(1e3 * (t2 - t0) / n, 3))
