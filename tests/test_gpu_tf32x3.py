"""GPU tests of the fp32 path on the tensor cores (3xTF32: kind::tf32 tcgen05 MMAs over hi/lo operand splits).  The
reference computes in fp32 (model/dim1/ABMIL.py:52-54 under train_ddp.py, no AMP), so the bound is the fp32 one: <= 1e-5
of the float64 oracle for the scores and every gradient, and agreement with the FFMA kernels it replaces."""
import numpy as np
import pytest
import torch

from oracle import mil_oracle as mo
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

D = 192


def _params(Lf, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    Wv = torch.randn(D, Lf, device="cuda", generator=gen) * 0.03
    Wu = torch.randn(D, Lf, device="cuda", generator=gen) * 0.03
    bv = torch.randn(D, device="cuda", generator=gen) * 0.1
    bu = torch.randn(D, device="cuda", generator=gen) * 0.1
    ww = torch.randn(D, device="cuda", generator=gen) * 0.3
    bw = torch.randn(1, device="cuda", generator=gen) * 0.1
    return gen, Wv, bv, Wu, bu, ww, bw


def _oracle_params(Wv, bv, Wu, bu, ww, bw):
    return {"attention_V.0.weight": Wv.double().cpu().numpy(), "attention_V.0.bias": bv.double().cpu().numpy(),
            "attention_U.0.weight": Wu.double().cpu().numpy(), "attention_U.0.bias": bu.double().cpu().numpy(),
            "attention_weights.weight": ww.view(1, -1).double().cpu().numpy(), "attention_weights.bias": bw.double().cpu().numpy()}


@pytest.mark.parametrize("n,Lf", [(1, 1024), (4096, 1024), (4097, 768), (5000, 512), (40000, 1024)])
def test_fp32_scores_on_tensor_cores_match_oracle(n, Lf, monkeypatch):
    from mil_b200 import functional as F
    gen, Wv, bv, Wu, bu, ww, bw = _params(Lf, 3)
    X = torch.randn(n, Lf, device="cuda", generator=gen)
    Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, X.dtype)
    s = F.gated_scores(X, Wcat, bcat, ww, bw)
    monkeypatch.setenv("MILB200_TF32X3", "0")
    s_ffma = F.gated_scores(X, Wcat, bcat, ww, bw)
    torch.cuda.synchronize()
    want, _, _ = mo.gated_scores(X.double().cpu().numpy(), *[t.double().cpu().numpy() for t in (Wv, bv, Wu, bu)],
                                 ww.view(1, -1).double().cpu().numpy(), bw.double().cpu().numpy())
    e_tc, e_ffma = rel_err(s.cpu().numpy(), want), rel_err(s_ffma.cpu().numpy(), want)
    assert e_tc <= 1e-5, (e_tc, e_ffma)
    assert rel_err(s.cpu().numpy(), s_ffma.cpu().numpy()) <= 1e-5


@pytest.mark.parametrize("lens,Lf", [([1], 1024), ([100, 31, 4000], 1024), ([5130, 2, 64], 768), ([3000, 5000, 33000], 1024)])
def test_fp32_gate_backward_on_tensor_cores_matches_oracle(lens, Lf, monkeypatch):
    from mil_b200 import functional as F
    gen, Wv, bv, Wu, bu, ww, bw = _params(Lf, 5)
    off = mo.offsets_from_lengths(np.asarray(lens))
    n = int(off[-1])
    X = torch.randn(n, Lf, device="cuda", generator=gen)
    offt = torch.from_numpy(off).cuda()
    dM = torch.randn(len(lens), Lf, device="cuda", generator=gen)
    Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, X.dtype)

    def run():
        s = F.gated_scores(X, Wcat, bcat, ww, bw)
        M, _, _, _ = F.segment_softmax_pool(X, s, offt)
        ds, _ = F.segment_softmax_pool_bwd(X, s, offt, dM, M, want_attn=False)
        _, dWcat, dbcat, dww, dbw = F.gated_scores_bwd(X, Wcat, bcat, ww, bw, ds, None, dM, offt, False)
        return [t.clone() for t in (dWcat, dbcat, dww, dbw)]

    got = run()
    monkeypatch.setenv("MILB200_TF32X3", "0")
    ffma = run()
    torch.cuda.synchronize()
    if n <= 10000:       # the float64 oracle in seconds
        g = mo.abmil_backward_csr(_oracle_params(Wv, bv, Wu, bu, ww, bw), X.double().cpu().numpy(), off,
                                  dM.double().cpu().numpy(), need_dx=False)
        dW = np.concatenate([g["attention_V.0.weight"], g["attention_U.0.weight"]], axis=0)
        db = np.concatenate([g["attention_V.0.bias"], g["attention_U.0.bias"]])
        errs = {}
        for name, i, want in (("dWcat", 0, dW), ("dbcat", 1, db), ("dww", 2, g["attention_weights.weight"].reshape(-1))):
            errs[name] = (rel_err(got[i].cpu().numpy().reshape(-1), want.reshape(-1)),
                          rel_err(ffma[i].cpu().numpy().reshape(-1), want.reshape(-1)))
        print("fp32 gate backward, error vs float64 oracle (3xTF32, FFMA):", errs)
        if n > 1:        # (a one-instance bag has ds = 0: every gradient is rounding noise around zero)
            for name, (e_tc, e_ffma) in errs.items():
                assert e_tc <= 1e-5, (name, e_tc, e_ffma)
    for a, b, name in zip(got, ffma, ("dWcat", "dbcat", "dww", "dbw")):
        if name == "dbw":
            assert abs(float(a) - float(b)) <= 1e-4
        else:
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5, name


def test_fp32_module_matches_reference_fixture_through_tensor_cores():
    """The nn.Module path with fp32 inputs (BASELINE configs[0]: 32 bags x 512 x 1024) runs the 3xTF32 kernels and stays
    within 1e-5 of the oracle for M and the parameter gradients."""
    import mil_b200
    lens = [512] * 10
    off = mo.offsets_from_lengths(np.asarray(lens))
    p = mo.procedural_state(mo.abmil_shapes(1024), 7)
    m = mil_b200.ABMIL(None, L=1024).cuda().eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    gen = torch.Generator(device="cuda").manual_seed(2)
    X = torch.randn(int(off[-1]), 1024, device="cuda", generator=gen)
    offt = torch.from_numpy(off).cuda()
    dM = torch.randn(len(lens), 1024, device="cuda", generator=gen)
    M = m.forward_csr(X, offt)
    (M * dM).sum().backward()
    torch.cuda.synchronize()
    Mo, _, _ = mo.abmil_forward_csr(p, X.double().cpu().numpy(), off)
    assert rel_err(M.detach().cpu().numpy(), Mo) <= 1e-5
    g = mo.abmil_backward_csr(p, X.double().cpu().numpy(), off, dM.double().cpu().numpy(), need_dx=False)
    for k, prm in m.state_dict(keep_vars=True).items():
        got, want = prm.grad.double().cpu().numpy().reshape(-1), np.asarray(g[k]).reshape(-1)
        if k == "attention_weights.bias":
            assert abs(got[0] - want[0]) <= 1e-5
        else:
            assert rel_err(got, want) <= 1e-5, k


@pytest.mark.parametrize("m,k,n,act,use_add", [(4096, 768, 512, "relu", False), (15592, 768, 512, None, False),
                                                (5000, 512, 256, "tanh", True), (4100, 512, 2048, "relu", False),
                                                (6000, 64, 128, None, False), (1000, 768, 512, "relu", False)])
def test_fp32_linear_on_tensor_cores(m, k, n, act, use_add, monkeypatch):
    """nn.Linear (+ReLU/Tanh, + the fused input add) with fp32 operands and m >= 4096 rows (one smaller case stays on FFMA): forward, dX, dW and dbias through
    the 3xTF32 kernels against float64 torch and against the FFMA kernels (aggregator.py:44,47,66; transformer.py:430-448)."""
    from mil_b200 import functional as F
    gen = torch.Generator(device="cuda").manual_seed(m + n)
    x = torch.randn(m, k, device="cuda", generator=gen)
    add = torch.randn(m, k, device="cuda", generator=gen) if use_add else None
    W = torch.randn(n, k, device="cuda", generator=gen) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=gen) * 0.1
    dy = torch.randn(m, n, device="cuda", generator=gen)

    def run():
        xs = [t.clone().requires_grad_(True) for t in ((x, add, W, b) if use_add else (x, W, b))]
        if use_add:
            y = F.linear(xs[0], xs[2], xs[3], act=act, add=xs[1])
        else:
            y = F.linear(xs[0], xs[1], xs[2], act=act)
        y.backward(dy)
        return [y.detach()] + [t.grad for t in xs]

    got = run()
    monkeypatch.setenv("MILB200_TF32X3", "0")
    ffma = run()
    xd = [t.double().clone().requires_grad_(True) for t in ((x, add, W, b) if use_add else (x, W, b))]
    xin = xd[0] + xd[1] if use_add else xd[0]
    Wd, bd = (xd[2], xd[3]) if use_add else (xd[1], xd[2])
    yd = xin @ Wd.t() + bd

    def check(res):
        # ReLU is discontinuous: a pre-activation within rounding of zero may land on either side, and one flipped unit
        # moves a whole row of dX by ~1 %.  The float64 backward therefore uses the sign pattern of the result under test.
        for t in xd:
            t.grad = None
        y2 = yd * (res[0] > 0).double() if act == "relu" else (torch.tanh(yd) if act == "tanh" else yd)
        y2.backward(dy.double(), retain_graph=True)
        want = [y2.detach()] + [t.grad.clone() for t in xd]
        return [rel_err(a.cpu().numpy(), w.cpu().numpy()) for a, w in zip(res, want)]

    torch.cuda.synchronize()
    e_tc, e_ffma = check(got), check(ffma)
    for i, e in enumerate(e_tc):
        assert e <= 1e-5, (i, e_tc, e_ffma)


@pytest.mark.parametrize("lens", [[3000, 1200, 5], [20000, 15000, 3000]])
def test_fp32_gate_input_gradient_on_tensor_cores(lens, monkeypatch):
    """dX of the gated pool with fp32 operands (the packed bag of the fusion path requires it): the GEMM term runs as 3xTF32
    with the pooling term a_i dM_b in its epilogue — also across the 32768-row chunking — against the FFMA kernels and,
    for the small case, the float64 oracle."""
    from mil_b200 import functional as F
    gen, Wv, bv, Wu, bu, ww, bw = _params(1024, 17)
    off = mo.offsets_from_lengths(np.asarray(lens))
    n = int(off[-1])
    X = torch.randn(n, 1024, device="cuda", generator=gen)
    offt = torch.from_numpy(off).cuda()
    dM = torch.randn(len(lens), 1024, device="cuda", generator=gen)
    Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, X.dtype)

    def run():
        s, act = F.gated_scores(X, Wcat, bcat, ww, bw, save=True)
        M, _, _, _ = F.segment_softmax_pool(X, s, offt)
        ds, attn = F.segment_softmax_pool_bwd(X, s, offt, dM, M, want_attn=True)
        dX, dWcat, _, _, _ = F.gated_scores_bwd(X, Wcat, bcat, ww, bw, ds, attn, dM, offt, True, gate_act=act)
        return dX.clone(), dWcat.clone()

    dX, dW = run()
    monkeypatch.setenv("MILB200_TF32X3", "0")
    dX0, dW0 = run()
    torch.cuda.synchronize()
    assert rel_err(dX.cpu().numpy(), dX0.cpu().numpy()) <= 1e-5
    assert rel_err(dW.cpu().numpy(), dW0.cpu().numpy()) <= 2e-5
    if n <= 10000:
        g = mo.abmil_backward_csr(_oracle_params(Wv, bv, Wu, bu, ww, bw), X.double().cpu().numpy(), off,
                                  dM.double().cpu().numpy(), need_dx=True)
        assert rel_err(dX.cpu().numpy(), g["x"]) <= 1e-5
