"""Smallest program that exercises the bench's step (config 2 workload) — the thing ncu wraps.
Usage: python tools/profile_step.py [steps] [bags]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mil_b200  # noqa: E402
from mil_b200.dp import AbmilTrainer  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bags = int(sys.argv[2]) if len(sys.argv) > 2 else 64
need_dx = len(sys.argv) > 3 and sys.argv[3] == "dx"
lengths = bench.bag_lengths(bags, 1234)
off = torch.zeros(bags + 1, dtype=torch.int32)
off[1:] = lengths.cumsum(0).to(torch.int32)
gen = torch.Generator(device="cuda").manual_seed(1234)
X = torch.randn(int(off[-1]), bench.L_FEAT, device="cuda", generator=gen).to(torch.bfloat16)
torch.manual_seed(1234)
m = mil_b200.ABMIL(None, L=bench.L_FEAT).cuda()
tr = AbmilTrainer(bench.L_FEAT, bench.D_GATE, torch.bfloat16, device="cuda", need_input_grad=need_dx)
tr.load_from(m)
offd = off.cuda()
for _ in range(steps):
    M = tr.step(X, offd)
torch.cuda.synchronize()
print("ok", float(M.sum()), mil_b200.launch_count())
