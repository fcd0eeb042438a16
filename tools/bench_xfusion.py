"""Timing of the collapsed fusion training step (FusionTrainer.step_bags, csrc/xfusion.cu): B patients per step,
N pathology rows each, 160 CT tokens, one text token, bf16.  Usage: python tools/bench_xfusion.py [N] [dtype] [B list]
Prints ms per step / per patient and library launches per patient.  MILB200_XF_ONCE=1: one step per B (for ncu)."""
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mil_b200  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 15592
    dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
    Bs = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "1,2,4,8".split(","))]
    once = os.environ.get("MILB200_XF_ONCE", "0") == "1"
    dev = torch.device("cuda", 0)
    ns = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                   aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
    torch.manual_seed(1234)
    m = mil_b200.get_model(ns).to(dev).eval()
    for B in Bs:
        tr = mil_b200.FusionTrainer(m, n_text_tokens=1, compute_dtype=dtype)
        ct = torch.randn(B, 160, 512, device=dev).to(dtype)
        xp = torch.randn(B * N, 768, device=dev).to(dtype)
        xt = (torch.randn(B, 512, device=dev) * 0.05).to(dtype)
        labels = torch.tensor([[0.0, 1.0]] * B, device=dev)
        lens = [N] * B
        if once:
            tr.step_bags(ct, xp, lens, xt, labels)
            torch.cuda.synchronize()
            continue
        for _ in range(5):
            tr.step_bags(ct, xp, lens, xt, labels)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = mil_b200.launch_count()
        reps = 20
        e0.record()
        for _ in range(reps):
            tr.step_bags(ct, xp, lens, xt, labels)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        import time
        t0 = time.perf_counter()
        for _ in range(reps):
            tr.step_bags(ct, xp, lens, xt, labels)
        host = (time.perf_counter() - t0) / reps * 1e3
        torch.cuda.synchronize()
        print(f"B={B} N={N} {sys.argv[2] if len(sys.argv) > 2 else 'bf16'}: {ms:.3f} ms/step  {ms / B:.3f} ms/patient  "
              f"{(mil_b200.launch_count() - l0) / reps / B:.1f} launches/patient  host enqueue {host:.3f} ms/step", flush=True)
        del tr


if __name__ == "__main__":
    main()
