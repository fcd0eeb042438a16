"""What bounds the gate GEMMs at full-chip load?  Loops one kernel for a few seconds while NVML samples SM clock and board
power (the bench's own clock sampler covers a 27 ms timed region with a handful of samples; this is the long version).
Usage: python tools/power_probe.py [seconds]"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mil_b200  # noqa: E402,F401
from mil_b200 import functional as F  # noqa: E402
import pynvml  # noqa: E402


def main():
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    g = torch.Generator().manual_seed(1234)
    lens = torch.randint(100, 20001, (64,), generator=g).numpy()
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(off[-1])
    gen = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, 1024, device="cuda", generator=gen).bfloat16()
    D = 192
    Wv = torch.randn(D, 1024, device="cuda", generator=gen) * 0.03
    Wu = torch.randn(D, 1024, device="cuda", generator=gen) * 0.03
    z = torch.zeros(D, device="cuda")
    ww = torch.randn(D, device="cuda", generator=gen) * 0.3
    bw = torch.zeros(1, device="cuda")
    dM = torch.randn(64, 1024, device="cuda", generator=gen)
    offt = torch.from_numpy(off).cuda()
    Wcat, bcat = F.pack_gate_weights(Wv, z, Wu, z, X.dtype)
    s, act = F.gated_scores(X, Wcat, bcat, ww, bw, save=True)
    M, _, _, _ = F.segment_softmax_pool(X, s, offt)
    ds, _ = F.segment_softmax_pool_bwd(X, s, offt, dM, M, want_attn=False)
    cases = {
        "gate_bwd dW (k_gemm_tn_gate)": lambda: F.gated_scores_bwd(X, Wcat, bcat, ww, bw, ds, None, dM, offt, False, gate_act=act),
        "gated_score_fwd (k_gemm_kmajor_2sm)": lambda: F.gated_scores(X, Wcat, bcat, ww, bw, save=True),
        "segment_softmax_pool_fwd (k_pool_fwd)": lambda: F.segment_softmax_pool(X, s, offt),
    }
    for name, fn in cases.items():
        samples, stop = [], threading.Event()

        def poll():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
                time.sleep(0.02)

        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        th = threading.Thread(target=poll)
        th.start()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        first = []
        reps = 0
        a.record()
        while time.perf_counter() - t0 < secs:
            for _ in range(50):
                fn()
            reps += 50
            if not first:
                b.record()
                torch.cuda.synchronize()
                first.append(a.elapsed_time(b) / 50)
                a.record()
                reps = 0
        b.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        ms_late = a.elapsed_time(b) / max(reps, 1)
        clk = np.array([c for c, _ in samples[len(samples) // 3:]])
        pw = np.array([p for _, p in samples[len(samples) // 3:]])
        print("%-42s first 50 calls %.4f ms/call, sustained %.4f ms/call | SM clock median %d MHz (min %d, max %d), power median "
              "%.0f W (max %.0f W), %d samples" % (name, first[0], ms_late, np.median(clk), clk.min(), clk.max(), np.median(pw),
                                                   pw.max(), len(clk)))
        time.sleep(1.0)


if __name__ == "__main__":
    main()
