"""Times the backward of the gated pool at BASELINE configs[1] size: two-kernel (pooling backward, then the gate
backward) against the mirrored single-pass backward (milb200_gated_pool_bwd).  CUDA events, median of N runs.
MILB200_TNG_LEAD=<k-blocks> changes how far the g warps run ahead."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mil_b200  # noqa: E402
from mil_b200 import functional as F  # noqa: E402


def main():
    g = torch.Generator().manual_seed(1234)
    lens = torch.randint(100, 20001, (64,), generator=g).numpy()
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(off[-1])
    gen = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, 1024, device="cuda", generator=gen).bfloat16()
    D = 192
    Wv = torch.randn(D, 1024, device="cuda", generator=gen) * 0.03
    Wu = torch.randn(D, 1024, device="cuda", generator=gen) * 0.03
    bv = torch.zeros(D, device="cuda")
    bu = torch.zeros(D, device="cuda")
    ww = torch.randn(D, device="cuda", generator=gen) * 0.3
    bw = torch.zeros(1, device="cuda")
    dM = torch.randn(64, 1024, device="cuda", generator=gen)
    offt = torch.from_numpy(off).cuda()
    Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, X.dtype)
    s, act = F.gated_scores(X, Wcat, bcat, ww, bw, save=True)
    M, _, _, _ = F.segment_softmax_pool(X, s, offt)

    def two():
        ds, _ = F.segment_softmax_pool_bwd(X, s, offt, dM, M, want_attn=False)
        return (ds,) + tuple(F.gated_scores_bwd(X, Wcat, bcat, ww, bw, ds, None, dM, offt, False, gate_act=act)[1:])

    def fused():
        return F.gated_pool_bwd(X, s, offt, dM, M, ww, act)

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts))

    if os.environ.get("MILB200_ONLY_FUSED"):
        for _ in range(6):
            fused()
        torch.cuda.synchronize()
        return
    r2 = two()
    rf = fused()
    torch.cuda.synchronize()
    for name, a, b in zip(("ds", "dWcat", "dbcat", "dww", "dbw"), rf, r2):
        err = float((a.double() - b.double()).abs().max() / b.double().abs().max())
        print(f"  {name}: rel err fused vs two-kernel {err:.2e}")
    print("instances", n, "lead", os.environ.get("MILB200_TNG_LEAD", "default"))
    print("two-kernel backward  ms (median, min): %.4f %.4f" % timeit(two))
    print("fused backward       ms (median, min): %.4f %.4f" % timeit(fused))
    if os.environ.get("MILB200_TRACE"):
        from mil_b200 import _lib as Lb
        tr = torch.zeros(16, dtype=torch.int64, device="cuda")
        Lb.check(Lb.lib().milb200_debug_trace(Lb.ptr(tr)), "trace")
        if os.environ.get("MILB200_TRACE") == "two":
            two()
        else:
            fused()
        torch.cuda.synchronize()
        Lb.lib().milb200_debug_trace(None)
        t = tr.cpu().tolist()
        print("CTA 0 trace (cycles): V,U producer chunks %d refetches %d aempty-wait %d | g warp 0: throttle %d rounds %d "
              "barrier+finish %d blocks %d | MMA: afull-wait %d bfull-wait %d loop %d stages %d" %
              (t[0], t[1], t[2], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11]))


if __name__ == "__main__":
    main()
