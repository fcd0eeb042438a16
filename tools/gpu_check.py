"""Kernel-by-kernel numerical check on a real B200 (developer tool, not part of the test-suite).
Each stage runs in its own subprocess under a timeout so a hung kernel cannot take the others down.
Usage: python tools/gpu_check.py [stage ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["pool_f32", "pool_bf16", "gate_f32", "gate_bf16_simt", "tc_score", "tc_linear", "tc_gate_bwd",
          "tc_linear_bwd", "abmil", "attention", "layernorm", "clip", "fusion"]


def rel(a, b):
    import torch
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def ref_gate(X, Wv, bv, Wu, bu, ww, bw):
    import torch
    Xd = X.double()
    V = torch.tanh(Xd @ Wv.double().t() + bv.double())
    U = torch.sigmoid(Xd @ Wu.double().t() + bu.double())
    return (V * U) @ ww.double().reshape(-1) + bw.double().reshape(-1)[0], V, U


def make_bags(lens, L, dtype, seed=0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = int(sum(lens))
    X = torch.randn(n, L, device="cuda", generator=g).to(dtype)
    off = torch.zeros(len(lens) + 1, dtype=torch.int32)
    off[1:] = torch.tensor(lens).cumsum(0)
    return X, off.cuda()


def stage_pool(dtype):
    import torch
    import mil_b200
    F = mil_b200.functional
    for L in (1024, 768, 512, 96 if dtype == torch.float32 else 64):
        for lens in ([1], [5, 300, 1, 129, 4000, 37, 128, 128, 256, 31], [20000, 100, 9000] * 10, [512] * 32):
            X, off = make_bags(lens, L, dtype, seed=L)
            n = X.shape[0]
            s = torch.randn(n, device="cuda") * 3
            M, Ml, am, lse = F.segment_softmax_pool(X, s, off, want_lowp=dtype != torch.float32)
            torch.cuda.synchronize()
            offc = off.cpu().tolist()
            Mr, amr, lser = [], [], []
            for b in range(len(lens)):
                sb = s[offc[b]:offc[b + 1]].double()
                a = torch.softmax(sb, 0)
                Mr.append(a @ X[offc[b]:offc[b + 1]].double())
                amr.append(int(sb.argmax()))
                lser.append(torch.logsumexp(sb, 0))
            Mr = torch.stack(Mr)
            e1 = rel(M, Mr)
            e2 = rel(lse, torch.stack(lser))
            okam = am.cpu().tolist() == amr
            dM = torch.randn(len(lens), L, device="cuda")
            ds, attn = F.segment_softmax_pool_bwd(X, s, off, dM, M, True)
            torch.cuda.synchronize()
            dsr, ar = [], []
            for b in range(len(lens)):
                sb = s[offc[b]:offc[b + 1]].double()
                a = torch.softmax(sb, 0)
                g = X[offc[b]:offc[b + 1]].double() @ dM[b].double()
                dsr.append(a * (g - (a * g).sum()))
                ar.append(a)
            e3 = rel(ds, torch.cat(dsr))
            e4 = rel(attn, torch.cat(ar))
            print(f"pool {dtype} L={L} B={len(lens)} n={n}: M {e1:.2e} lse {e2:.2e} argmax {okam} ds {e3:.2e} attn {e4:.2e}",
                  flush=True)


def stage_gate(dtype, with_dx=True, Ls=(1024, 512, 768), n_list=(1, 300, 5000)):
    import torch
    import mil_b200
    F = mil_b200.functional
    D = 192
    for L in Ls:
        for n in n_list:
            lens = [n] if n < 100 else [n // 3, n - n // 3 - 7, 7]
            X, off = make_bags(lens, L, dtype, seed=n)
            g = torch.Generator(device="cuda").manual_seed(1)
            k = 1.0 / L ** 0.5
            Wv = (torch.rand(D, L, device="cuda", generator=g) * 2 - 1) * k
            Wu = (torch.rand(D, L, device="cuda", generator=g) * 2 - 1) * k
            bv = (torch.rand(D, device="cuda", generator=g) * 2 - 1) * k
            bu = (torch.rand(D, device="cuda", generator=g) * 2 - 1) * k
            ww = (torch.rand(1, D, device="cuda", generator=g) * 2 - 1) / D ** 0.5
            bw = torch.rand(1, device="cuda", generator=g)
            Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, dtype)
            s, act = F.gated_scores(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw, save=True)
            s2 = F.gated_scores(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw)
            torch.cuda.synchronize()
            assert torch.equal(s, s2), "saving the gate activations changed the scores"
            Wvq, Wuq = Wv.to(dtype).float(), Wu.to(dtype).float()
            sr, V, U = ref_gate(X, Wvq, bv, Wuq, bu, ww, bw)
            e_s = rel(s, sr)
            ds = torch.randn(n, device="cuda") / n ** 0.5
            attn = torch.rand(n, device="cuda")
            dM = torch.randn(len(lens), L, device="cuda")
            dX, dWcat, dbcat, dww, dbw = F.gated_scores_bwd(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw, ds, attn,
                                                            dM, off, with_dx)
            torch.cuda.synchronize()
            dG = ds.double()[:, None] * ww.double().reshape(1, -1)
            dVp = dG * U * (1 - V * V)
            dUp = dG * V * U * (1 - U)
            dWr = torch.cat([dVp.t() @ X.double(), dUp.t() @ X.double()])
            dbr = torch.cat([dVp.sum(0), dUp.sum(0)])
            dwwr = ds.double() @ (V * U)
            if act is not None:   # saved-activation backward vs the recompute backward
                dX2, dW2, db2, dww2, dbw2 = (t.clone() if t is not None else None for t in
                                              F.gated_scores_bwd(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw, ds, attn,
                                                                 dM, off, with_dx, gate_act=act))
                torch.cuda.synchronize()
                print(f"   saved-vs-recompute: dW {rel(dW2, dWcat):.2e} db {rel(db2, dbcat):.2e} dww {rel(dww2, dww):.2e} "
                      f"dbw {abs(float(dbw2) - float(dbw)):.1e} dX {rel(dX2, dX) if with_dx else -1:.2e}", flush=True)
                # no input gradient wanted: the fused dW GEMM builds dZ on the fly (k_gemm_tn_gate)
                _, dW3, db3, dww3, dbw3 = (t.clone() if t is not None else None for t in
                                           F.gated_scores_bwd(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw, ds, None, dM,
                                                              off, False, gate_act=act))
                torch.cuda.synchronize()
                print(f"   fused-dW-vs-recompute: dW {rel(dW3, dWcat):.2e} db {rel(db3, dbcat):.2e} dww {rel(dww3, dww):.2e} "
                      f"dbw {abs(float(dbw3) - float(dbw)):.1e}   vs oracle: dW {rel(dW3, dWr):.2e} db {rel(db3, dbr):.2e} "
                      f"dww {rel(dww3, dwwr):.2e}", flush=True)
                dX, dWcat, dbcat, dww, dbw = dX2, dW2, db2, dww2, dbw2
            dG = ds.double()[:, None] * ww.double().reshape(1, -1)
            dVp = dG * U * (1 - V * V)
            dUp = dG * V * U * (1 - U)
            dWr = torch.cat([dVp.t() @ X.double(), dUp.t() @ X.double()])
            dbr = torch.cat([dVp.sum(0), dUp.sum(0)])
            dwwr = ds.double() @ (V * U)
            e_w, e_b, e_ww = rel(dWcat, dWr), rel(dbcat, dbr), rel(dww, dwwr)
            e_bw = abs(float(dbw) - float(ds.double().sum()))
            e_x = -1.0
            if with_dx:
                bag = torch.bucketize(torch.arange(n, device="cuda"), off[1:].long(), right=True)
                dXr = attn.double()[:, None] * dM.double()[bag] + dVp @ Wvq.double() + dUp @ Wuq.double()
                e_x = rel(dX, dXr)
            print(f"gate {dtype} L={L} n={n}: s {e_s:.2e} dW {e_w:.2e} db {e_b:.2e} dww {e_ww:.2e} dbw {e_bw:.1e} dX {e_x:.2e}",
                  flush=True)


def stage_tc_score():
    import torch
    import mil_b200
    F = mil_b200.functional
    D = 192
    for L in (1024, 512, 768, 64):
        for n in (128, 1, 300, 5000, 40000):
            X, off = make_bags([n], L, torch.bfloat16, seed=n)
            g = torch.Generator(device="cuda").manual_seed(1)
            k = 1.0 / L ** 0.5
            Wv = (torch.rand(D, L, device="cuda", generator=g) * 2 - 1) * k
            Wu = (torch.rand(D, L, device="cuda", generator=g) * 2 - 1) * k
            bv = (torch.rand(D, device="cuda", generator=g) * 2 - 1) * k
            bu = (torch.rand(D, device="cuda", generator=g) * 2 - 1) * k
            ww = (torch.rand(1, D, device="cuda", generator=g) * 2 - 1) / D ** 0.5
            bw = torch.rand(1, device="cuda", generator=g)
            Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, torch.bfloat16)
            s = F.gated_scores(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw)
            torch.cuda.synchronize()
            sr, V, U = ref_gate(X, Wv.bfloat16().float(), bv, Wu.bfloat16().float(), bu, ww, bw)
            print(f"tc_score L={L} n={n}: s rel {rel(s, sr):.2e}  (s[0..3]={s[:3].tolist()} ref={sr[:3].tolist()})", flush=True)


def _linear(x, add, W, b, act, back=False, Y=None, dY=None, need_dx=True):
    import torch
    from mil_b200 import _lib as Lb
    m, k = x.shape
    n = W.shape[0]
    code = Lb.dtype_code(x)
    nb = Lb.lib().milb200_linear_workspace_bytes(m, n, k, code, 1)
    ws = Lb.workspace(nb, x.device)
    if not back:
        Y = torch.empty(m, n, dtype=x.dtype, device=x.device)
        Lb.check(Lb.lib().milb200_linear_fwd(Lb.ptr(x), Lb.ptr(add), Lb.ptr(W), Lb.ptr(b), Lb.ptr(Y), m, n, k, act, code,
                                             Lb.ptr(ws), ws.numel(), Lb.stream_ptr()), "linear_fwd")
        return Y
    dX = torch.empty_like(x) if need_dx else None
    dW = torch.empty(n, k, dtype=torch.float32, device=x.device)
    db = torch.empty(n, dtype=torch.float32, device=x.device)
    Lb.check(Lb.lib().milb200_linear_bwd(Lb.ptr(x), Lb.ptr(add), Lb.ptr(W), Lb.ptr(Y), Lb.ptr(dY), Lb.ptr(dX), Lb.ptr(dW),
                                         Lb.ptr(db), m, n, k, act, code, 0, Lb.ptr(ws), ws.numel(), Lb.stream_ptr()),
             "linear_bwd")
    return dX, dW, db


def stage_linear(dtype, back):
    import torch
    acts = {0: lambda t: t, 1: torch.tanh, 2: torch.relu}
    for (m, n, k) in [(300, 512, 768), (1, 512, 512), (5000, 256, 512), (10, 2048, 512), (10, 512, 2048), (160, 512, 512),
                      (20000, 512, 768), (33, 2, 512), (64, 384, 1536)]:
        for act in (0, 1, 2):
            g = torch.Generator(device="cuda").manual_seed(m + n)
            x = torch.randn(m, k, device="cuda", generator=g).to(dtype)
            add = (torch.randn(m, k, device="cuda", generator=g).to(dtype)) if act == 1 else None
            W = ((torch.rand(n, k, device="cuda", generator=g) * 2 - 1) / k ** 0.5).to(dtype)
            b = (torch.rand(n, device="cuda", generator=g) * 2 - 1) / k ** 0.5
            Y = _linear(x, add, W, b, act)
            torch.cuda.synchronize()
            xin = x.double() + (add.double() if add is not None else 0)
            pre = xin @ W.double().t() + b.double()
            Yr = acts[act](pre)
            msg = f"linear {dtype} m={m} n={n} k={k} act={act}: Y {rel(Y, Yr):.2e}"
            if back:
                dY = torch.randn(m, n, device="cuda", generator=g).to(dtype)
                dX, dW, db = _linear(x, add, W, b, act, True, Y, dY)
                torch.cuda.synchronize()
                Yd = Y.double()
                dpre = dY.double() * ({0: 1.0, 1: 1 - Yd * Yd, 2: (Yd > 0).double()}[act])
                msg += f" dX {rel(dX, dpre @ W.double()):.2e} dW {rel(dW, dpre.t() @ xin):.2e} db {rel(db, dpre.sum(0)):.2e}"
            print(msg, flush=True)


def stage_abmil():
    import numpy as np
    import torch
    import mil_b200
    from oracle import mil_oracle as mo
    for dtype, L, tol in ((torch.float32, 1024, 1e-5), (torch.bfloat16, 1024, 1e-2), (torch.float32, 96, 1e-5),
                          (torch.bfloat16, 512, 1e-2)):
        p = mo.procedural_state(mo.abmil_shapes(L), 3)
        lens = mo.ragged_lengths(6, 1, 700, 5)
        off = mo.offsets_from_lengths(lens)
        X = np.random.RandomState(1).standard_normal((int(off[-1]), L)).astype(np.float32)
        Xt = torch.from_numpy(X).cuda().to(dtype)
        m = mil_b200.ABMIL(None, L=L).cuda().eval()
        m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        Xt.requires_grad_(True)
        M = m.forward_csr(Xt, torch.from_numpy(off).cuda())
        dM = np.random.RandomState(2).standard_normal(M.shape).astype(np.float32)
        (M.float() * torch.from_numpy(dM).cuda()).sum().backward()
        torch.cuda.synchronize()
        pq = {k: (torch.from_numpy(v).to(dtype).float().numpy() if k.endswith("0.weight") else v) for k, v in p.items()}
        Xq = Xt.detach().float().cpu().numpy()
        Mr, sr, amr = mo.abmil_forward_csr(pq, Xq, off)
        gr = mo.abmil_backward_csr(pq, Xq, off, dM)
        r = lambda a, b: float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
        out = [f"M {r(M.detach().float().cpu().numpy(), Mr):.2e}", f"s {r(m.last_scores.cpu().numpy(), sr):.2e}",
               f"argmax {m.last_argmax.cpu().numpy().tolist() == amr.tolist()}",
               f"dX {r(Xt.grad.float().cpu().numpy(), gr['x']):.2e}"]
        for k, prm in m.state_dict(keep_vars=True).items():
            if k == "attention_weights.bias":
                continue
            out.append(f"{k.split('.')[0][-1]}{k.split('.')[-1][0]} {r(prm.grad.float().cpu().numpy(), gr[k]):.2e}")
        print(f"abmil {dtype} L={L} tol={tol}: " + " ".join(out), flush=True)


def stage_attention():
    import torch
    from mil_b200 import functional as F
    for dtype in (torch.float32, torch.bfloat16):
        for (nq, nk, H, c) in [(10, 5000, 8, 32), (1, 160, 8, 32), (1, 20000, 8, 32), (16, 333, 8, 32), (5000, 10, 8, 32),
                               (160, 1, 8, 32), (20000, 1, 8, 32), (10, 10, 8, 64), (1, 1, 8, 64), (3, 3, 8, 64), (7, 17, 8, 32)]:
            g = torch.Generator(device="cuda").manual_seed(nq * 31 + nk)
            q = torch.randn(nq, H * c, device="cuda", generator=g).to(dtype).requires_grad_(True)
            k = torch.randn(nk, H * c, device="cuda", generator=g).to(dtype).requires_grad_(True)
            v = torch.randn(nk, H * c, device="cuda", generator=g).to(dtype).requires_grad_(True)
            do = torch.randn(nq, H * c, device="cuda", generator=g).to(dtype)
            o = F.attention_core(q, k, v, H)
            (o.float() * do.float()).sum().backward()
            torch.cuda.synchronize()
            qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
            qh = qd.view(nq, H, c).transpose(0, 1); kh = kd.view(nk, H, c).transpose(0, 1); vh = vd.view(nk, H, c).transpose(0, 1)
            a = torch.softmax(qh @ kh.transpose(1, 2) / c ** 0.5, -1)
            orf = (a @ vh).transpose(0, 1).reshape(nq, H * c)
            (orf * do.double()).sum().backward()
            print(f"attention {dtype} nq={nq} nk={nk} c={c}: O {rel(o, orf):.2e} dQ {rel(q.grad, qd.grad):.2e} "
                  f"dK {rel(k.grad, kd.grad):.2e} dV {rel(v.grad, vd.grad):.2e}", flush=True)


def stage_layernorm():
    import torch
    from mil_b200 import functional as F
    for dtype in (torch.float32, torch.bfloat16):
        for (m, n, res) in [(1, 512, False), (10, 512, True), (5000, 512, True), (20000, 512, False), (77, 64, True), (3, 768, True)]:
            g = torch.Generator(device="cuda").manual_seed(m + n)
            x = torch.randn(m, n, device="cuda", generator=g).to(dtype).requires_grad_(True)
            r = torch.randn(m, n, device="cuda", generator=g).to(dtype).requires_grad_(True) if res else None
            ga = (torch.rand(n, device="cuda", generator=g) + 0.5).requires_grad_(True)
            be = torch.randn(n, device="cuda", generator=g).requires_grad_(True)
            dy = torch.randn(m, n, device="cuda", generator=g).to(dtype)
            y = F.layernorm(x, ga, be, residual=r)
            (y.float() * dy.float()).sum().backward()
            torch.cuda.synchronize()
            xd = x.detach().double().requires_grad_(True)
            rd = r.detach().double().requires_grad_(True) if res else None
            gd, bd = ga.detach().double().requires_grad_(True), be.detach().double().requires_grad_(True)
            yr = torch.nn.functional.layer_norm(xd + (rd if res else 0), (n,), gd, bd, 1e-5)
            (yr * dy.double()).sum().backward()
            msg = f"layernorm {dtype} m={m} n={n} res={res}: Y {rel(y, yr):.2e} dX {rel(x.grad, xd.grad):.2e} dg {rel(ga.grad, gd.grad):.2e} db {rel(be.grad, bd.grad):.2e}"
            if res:
                msg += f" dR {rel(r.grad, rd.grad):.2e}"
            print(msg, flush=True)


def stage_clip():
    import math
    import numpy as np
    import torch
    from mil_b200 import functional as F
    from oracle import mil_oracle as mo
    for dtype in (torch.float32, torch.bfloat16):
        for (bi, bt, d) in [(24, 24, 512), (64, 64, 512), (5, 9, 512), (1, 1, 512), (300, 300, 512)]:
            g = torch.Generator(device="cuda").manual_seed(bi)
            img = torch.randn(bi, d, device="cuda", generator=g).to(dtype).requires_grad_(True)
            txt = torch.randn(bt, d, device="cuda", generator=g).to(dtype).requires_grad_(True)
            ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
            dli = torch.randn(bi, bt, device="cuda", generator=g)
            dlt = torch.randn(bt, bi, device="cuda", generator=g)
            li, lt = F.clip_logits(img, txt, ls)
            ((li * dli).sum() + (lt * dlt).sum()).backward()
            torch.cuda.synchronize()
            i_np, t_np = img.detach().float().cpu().numpy(), txt.detach().float().cpu().numpy()
            lir, ltr = mo.clip_cosine_logits(i_np, t_np, float(ls))
            gi, gt, gs = mo.clip_cosine_logits_bwd(i_np, t_np, float(ls), dli.cpu().numpy(), dlt.cpu().numpy())
            r = lambda a, b: float(np.abs(np.asarray(a.detach().float().cpu().numpy(), dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
            print(f"clip_logits {dtype} bi={bi} bt={bt}: li {r(li, lir):.2e} lt {r(lt, ltr):.2e} dI {r(img.grad, gi):.2e} dT {r(txt.grad, gt):.2e} "
                  f"dscale {abs(float(ls.grad) - gs) / max(abs(gs), 1e-30):.2e}", flush=True)
        for (b, I, d) in [(6, 9, 512), (64, 9, 512), (1, 9, 512), (200, 3, 512)]:
            g = torch.Generator(device="cuda").manual_seed(b)
            out = (torch.randn(b, d, device="cuda", generator=g) * 0.3).to(dtype).requires_grad_(True)
            feat = (torch.randn(b, I, d, device="cuda", generator=g) * 0.3).to(dtype)
            loss, logits = F.cliploss_v1(out, feat)
            (loss * 1.7).backward()
            torch.cuda.synchronize()
            o_np, f_np = out.detach().float().cpu().numpy(), feat.float().cpu().numpy()
            lr_, lg = mo.cliploss_v1(o_np, f_np)
            dr = mo.cliploss_v1_bwd(o_np, f_np)
            print(f"cliploss {dtype} b={b} I={I}: loss {abs(float(loss) - lr_) / max(abs(lr_), 1e-30):.2e} logits "
                  f"{float(np.abs(logits.cpu().numpy() - lg).max() / max(np.abs(lg).max(), 1e-30)):.2e} dout "
                  f"{float(np.abs(out.grad.float().cpu().numpy() / 1.7 - dr).max() / max(np.abs(dr).max(), 1e-30)):.2e}", flush=True)
        for (n, d) in [(1, 512), (10, 512), (64, 512)]:
            g = torch.Generator(device="cuda").manual_seed(n)
            a = torch.randn(n, d, device="cuda", generator=g).to(dtype).requires_grad_(True)
            b = torch.randn(n, d, device="cuda", generator=g).to(dtype).requires_grad_(True)
            loss = F.cosine_embedding_loss(a, b)
            loss.backward()
            ad, bd = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
            lr_ = torch.nn.CosineEmbeddingLoss()(ad, bd, torch.ones(n, device="cuda", dtype=torch.float64))
            lr_.backward()
            print(f"cosine {dtype} n={n}: loss {abs(float(loss) - float(lr_)):.2e} da {rel(a.grad, ad.grad):.2e} db {rel(b.grad, bd.grad):.2e}", flush=True)
    z = torch.randn(7, 2, device="cuda", requires_grad=True)
    t = torch.randint(0, 2, (7, 2), device="cuda").float()
    loss, prob = F.sigmoid_bce(z, t)
    loss.backward()
    zd = z.detach().double().requires_grad_(True)
    lr_ = torch.nn.BCELoss()(torch.sigmoid(zd), t.double())
    lr_.backward()
    print(f"bce: loss {abs(float(loss) - float(lr_)):.2e} dz {rel(z.grad, zd.grad):.2e} prob {rel(prob, torch.sigmoid(zd)):.2e}", flush=True)


def _sd_np(module):
    return {k: v.detach().float().cpu().numpy() for k, v in module.state_dict().items()}


def stage_fusion():
    """Module level: TwoWayAttentionBlock / TwoWayTransformer / aggregator at the real dims vs the float64 oracle."""
    import numpy as np
    import torch
    import mil_b200
    from argparse import Namespace
    from oracle import fusion_oracle as fo
    from oracle import mil_oracle as mo
    from tests.test_oracle_golden import aggregator_shapes, transformer_shapes
    args = Namespace(modality=["CT", "pathology"], model_CT="resnetMC3_18", model_pathology="ABMIL", model_CI="none",
                     aggregator="ABMIL", num_classes=2, alignment_base="none", clinical_features=list("abcdefghi"))
    for dtype in (torch.float32, torch.bfloat16):
        for T, N in ((1, 300), (10, 1000), (3, 5000)):
            sdn = mo.procedural_state(transformer_shapes(512, 2048), 7)
            m = mil_b200.TwoWayTransformer(args=args, depth=2, embedding_dim=512, num_heads=8, mlp_dim=2048).cuda()
            m.load_state_dict({k: torch.from_numpy(v) for k, v in sdn.items()})
            g = torch.Generator(device="cuda").manual_seed(T * 1000 + N)
            img = torch.randn(1, N, 512, device="cuda", generator=g).to(dtype).requires_grad_(True)
            pe = torch.randn(1, N, 512, device="cuda", generator=g).to(dtype)
            pt = (torch.randn(1, T, 512, device="cuda", generator=g) * 0.5).to(dtype).requires_grad_(True)
            dq = torch.randn(1, T, 512, device="cuda", generator=g)
            dk = torch.randn(1, N, 512, device="cuda", generator=g)
            oq, ok = m(img, pe, pt)
            ((oq.float() * dq).sum() + (ok.float() * dk).sum()).backward()
            torch.cuda.synchronize()
            quant = (lambda a: torch.from_numpy(a).to(dtype).double()) if dtype != torch.float32 else (lambda a: torch.from_numpy(a).double())
            sd = {k: (quant(v) if v.ndim == 2 else torch.from_numpy(v).double()).requires_grad_(True) for k, v in sdn.items()}
            sd = {"t." + k: v for k, v in sd.items()}
            imgd = img.detach().double().cpu().requires_grad_(True)
            ptd = pt.detach().double().cpu().requires_grad_(True)
            roq, rok = fo.two_way_transformer(sd, "t", imgd, pe.double().cpu(), ptd)
            ((roq * dq.double().cpu()).sum() + (rok * dk.double().cpu()).sum()).backward()
            worst, wname = 0.0, ""
            for k_, p_ in m.named_parameters():
                rg = sd["t." + k_].grad
                if rg is None or p_.grad is None:
                    continue
                if k_.endswith("k_proj.bias"):
                    continue
                e = rel(p_.grad.cpu(), rg)
                if e > worst:
                    worst, wname = e, k_
            print(f"twoway {dtype} T={T} N={N}: oq {rel(oq.cpu(), roq):.2e} ok {rel(ok.cpu(), rok):.2e} dimg {rel(img.grad.cpu(), imgd.grad):.2e} "
                  f"dpt {rel(pt.grad.cpu(), ptd.grad):.2e} worst param grad {worst:.2e} ({wname})", flush=True)
    # full aggregator vs the reference fixtures
    from tests.helpers import load_golden, rnd, digest
    for name in ("aggregator_T1_N70", "aggregator_T10_N45"):
        fx = load_golden(name)
        seed, T, N = int(fx["seed"]), int(fx["T"]), int(fx["N"])
        m = mil_b200.get_model(args).cuda().eval()
        sdn = mo.procedural_state(aggregator_shapes(), seed)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sdn.items()})
        x_ct = torch.from_numpy(rnd(seed + 100, 1, 512, 160, 1, 2)).cuda().requires_grad_(True)
        x_p = torch.from_numpy(rnd(seed + 200, 1, N, 768)).cuda().requires_grad_(True)
        x_t = torch.from_numpy(rnd(seed + 300, 1, T, 512, scale=0.05)).cuda()
        prob, ct2ci, pth2ci = m([x_ct, x_p], x_t)
        label = torch.tensor([[0.0, 1.0]], device="cuda")
        loss = torch.nn.BCELoss()(prob, label) + mil_b200.clip_loss.cosine_embedding_loss(ct2ci.squeeze(0), pth2ci.squeeze(0))
        loss.backward()
        torch.cuda.synchronize()
        r = lambda a, b: float(np.abs(a.detach().cpu().numpy().astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30))
        worst, wname = 0.0, ""
        for k_, p_ in m.named_parameters():
            ref = fx.get("g:" + k_)
            if ref is None or p_.grad is None or k_.endswith("k_proj.bias") or k_.endswith("attention_weights.bias"):
                continue
            d = digest(p_.grad.cpu().numpy())
            e = abs(d[1] - ref[1]) / max(ref[1], 1e-30)
            if e > worst:
                worst, wname = e, k_
        print(f"{name}: prob {r(prob, fx['prob']):.2e} ct2ci {r(ct2ci, fx['ct2ci']):.2e} pth2ci {r(pth2ci, fx['pth2ci']):.2e} "
              f"loss {abs(float(loss) - float(fx['loss'])):.2e} worst sumsq-digest rel {worst:.2e} ({wname}) "
              f"dx_p sumsq {abs(digest(x_p.grad.cpu().numpy())[1] - fx['dx_p'][1]) / fx['dx_p'][1]:.2e}", flush=True)


def run_stage(name):
    import torch
    if name == "pool_f32":
        stage_pool(torch.float32)
    elif name == "pool_bf16":
        stage_pool(torch.bfloat16)
    elif name == "gate_f32":
        stage_gate(torch.float32)
    elif name == "gate_bf16_simt":
        stage_gate(torch.bfloat16)
    elif name == "tc_score":
        stage_tc_score()
    elif name == "tc_linear":
        stage_linear(torch.bfloat16, False)
        stage_linear(torch.float32, False)
    elif name == "tc_gate_bwd":
        stage_gate(torch.bfloat16, n_list=(1, 300, 5000, 40000))
    elif name == "tc_linear_bwd":
        stage_linear(torch.bfloat16, True)
        stage_linear(torch.float32, True)
    elif name == "abmil":
        stage_abmil()
    elif name == "attention":
        stage_attention()
    elif name == "layernorm":
        stage_layernorm()
    elif name == "clip":
        stage_clip()
    elif name == "fusion":
        stage_fusion()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        run_stage(sys.argv[2])
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for st in stages:
        env = dict(os.environ)
        if st.endswith("_simt"):
            env["MILB200_FORCE_SIMT"] = "1"
        t0 = time.time()
        print(f"===== {st}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", st], env=env, timeout=240,
                               capture_output=True, text=True)
            print(r.stdout[-6000:], flush=True)
            if r.returncode != 0:
                print(f"[{st}] exit {r.returncode}\n{r.stderr[-3000:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"[{st}] TIMEOUT\n{(e.stdout or b'')[-3000:]}", flush=True)
        print(f"----- {st} {time.time() - t0:.1f}s", flush=True)
