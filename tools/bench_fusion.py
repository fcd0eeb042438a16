"""Secondary measurement (BASELINE configs[2]/[3], not the driver's bench line): fwd+bwd of the full multimodal
aggregator (aggregator.py CT+pathology branch, ABMIL aggregator) and of aggregator_clip on WSI-scale bags, one bag per
call as the reference trains (train_ddp.py:75).  Prints one JSON line per case: ms per bag, bags/s, libmilb200 launches
per bag, and the top kernels by time share (CUDA events around the public module call).
Usage: python tools/bench_fusion.py [--cpu]   (--cpu also times the oracle's torch restatement on the host cores)"""
import json
import os
import sys
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mil_b200  # noqa: E402

ARGS = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                 aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = mil_b200.launch_count()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (mil_b200.launch_count() - l0) / reps


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    out = []
    for dtype in (torch.float32, torch.bfloat16):
        m = mil_b200.get_model(ARGS).to(dev).to(dtype).train(False)
        for T, N in ((1, 1000), (1, 15592), (10, 15592)):
            x_ct = torch.randn(1, 512, 160, 1, 1, device=dev, dtype=dtype)
            x_p = torch.randn(1, N, 768, device=dev, dtype=dtype)
            x_t = (torch.randn(1, T, 512, device=dev) * 0.05).to(dtype)
            label = torch.tensor([[0.0, 1.0]], device=dev)

            plist = list(m.parameters())

            def step():
                for p in plist:
                    p.grad = None
                prob, a, b = m([x_ct, x_p], x_t)
                loss = torch.nn.functional.binary_cross_entropy(prob.float(), label) + \
                    mil_b200.clip_loss.cosine_embedding_loss(a.squeeze(0), b.squeeze(0)).float()
                loss.backward()

            ms, launches = timed(step, 10)

            def fwd_only():
                with torch.no_grad():
                    m([x_ct, x_p], x_t)

            ms_f, l_f = timed(fwd_only, 10)
            # algorithmic image-side GEMM flops: fc_pathology 2*N*768*512 + 10 projections 2*N*512*256 (fwd), x3 fwd+bwd
            flops = 3 * (2 * N * 768 * 512 + 10 * 2 * N * 512 * 256) + 3 * 10 * 2 * 160 * 512 * 256
            out.append({"case": "aggregator CT+pathology fwd+bwd", "dtype": str(dtype).split(".")[-1], "T": T, "N": N,
                        "ms_per_bag": ms, "bags_per_s": 1e3 / ms, "launches_per_bag": launches, "fwd_only_ms": ms_f,
                        "fwd_only_launches": l_f, "image_side_gemm_tflops": flops / ms / 1e9})
            print(json.dumps(out[-1]), flush=True)
    # aggregator_clip, batched CSR entry: 64 bags of 100..15592 x 768 + CLIPloss_v1 (I = 9) + CLIP logits
    args = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", num_classes=2)
    for dtype in (torch.float32, torch.bfloat16):
        m = mil_b200.model.utils_clip.get_model(args).to(dev).to(dtype).train(False)
        g = torch.Generator().manual_seed(1234)
        lens = torch.randint(100, 15593, (64,), generator=g)
        off = torch.zeros(65, dtype=torch.int32)
        off[1:] = lens.cumsum(0)
        X = torch.randn(int(off[-1]), 768, device=dev, dtype=dtype)
        offd = off.to(dev)
        x_ct = torch.randn(64, 512, device=dev, dtype=dtype)
        feats = (torch.randn(64, 9, 512, device=dev) * 0.3).to(dtype)
        crit = mil_b200.CLIPloss_v1(Namespace(clinical_features=list("abcdefghi")))
        head = mil_b200.CLIPLogits().to(dev)

        def step():
            for p in m.parameters():
                p.grad = None
            a, b, prob = m.forward_csr(x_ct, X, offd)
            li, lt = head(a, b)
            loss = crit(b, feats) + li.diagonal().mean() * 1e-3 + prob.float().mean()
            loss.backward()

        ms, launches = timed(step, 10)
        nbytes = X.numel() * X.element_size()
        out.append({"case": "aggregator_clip forward_csr + CLIPloss_v1 + CLIP logits fwd+bwd", "dtype": str(dtype).split(".")[-1],
                    "bags": 64, "instances": int(off[-1]), "ms_per_step": ms, "bags_per_s": 64e3 / ms,
                    "launches_per_step": launches, "x_passes_equiv_gbs": 4 * nbytes / ms / 1e6})
        print(json.dumps(out[-1]), flush=True)
    if "--cpu" in sys.argv:
        from oracle import fusion_oracle as fo
        from oracle import mil_oracle as mo
        from tests.test_oracle_golden import aggregator_shapes
        torch.set_num_threads(os.cpu_count() or 1)
        sd = fo.to_torch(mo.procedural_state(aggregator_shapes(), 1), dtype=torch.float32, requires_grad=True)
        for T, N in ((1, 1000), (1, 15592)):
            x_ct = torch.randn(1, 512, 160, 1, 1)
            x_p = torch.randn(1, N, 768)
            x_t = torch.randn(1, T, 512) * 0.05
            ts = []
            for _ in range(3):
                for v in sd.values():
                    v.grad = None
                t0 = time.perf_counter()
                prob, a, b = fo.aggregator_fusion_forward(sd, x_ct, x_p, x_t)
                (prob.sum() + (a * b).sum()).backward()
                ts.append(time.perf_counter() - t0)
            print(json.dumps({"case": "oracle port (CPU, fp32) aggregator fwd+bwd", "T": T, "N": N, "ms_per_bag": 1e3 * min(ts),
                              "cores": os.cpu_count()}), flush=True)


if __name__ == "__main__":
    main()
