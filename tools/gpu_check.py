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
          "tc_linear_bwd", "abmil"]


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
            s = F.gated_scores(X, Wcat, bcat, ww.reshape(-1).contiguous(), bw)
            torch.cuda.synchronize()
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
