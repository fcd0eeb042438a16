"""GPU tests of the mirrored single-pass backward (milb200_gated_pool_bwd): the pooling backward computed inside the dWcat
tensor-core kernel must give what the two-kernel backward gives (milb200_segment_softmax_pool_bwd followed by
milb200_gated_score_bwd) — on ragged bags whose boundaries fall anywhere inside the kernel's 32-row k-blocks — and what
the float64 oracle gives (autograd of model/dim1/ABMIL.py:52-59)."""
import numpy as np
import pytest
import torch

from oracle import mil_oracle as mo
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

L_FEAT, D = 1024, 192


def _setup(lens, seed, Lf=L_FEAT):
    from mil_b200 import functional as F
    off = mo.offsets_from_lengths(np.asarray(lens))
    n = int(off[-1])
    gen = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(n, Lf, device="cuda", generator=gen).bfloat16()
    Wv = torch.randn(D, Lf, device="cuda", generator=gen) * 0.03
    Wu = torch.randn(D, Lf, device="cuda", generator=gen) * 0.03
    bv = torch.randn(D, device="cuda", generator=gen) * 0.1
    bu = torch.randn(D, device="cuda", generator=gen) * 0.1
    ww = torch.randn(D, device="cuda", generator=gen) * 0.3
    bw = torch.randn(1, device="cuda", generator=gen) * 0.1
    dM = torch.randn(len(lens), Lf, device="cuda", generator=gen)
    offt = torch.from_numpy(off).cuda()
    Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bu, X.dtype)
    s, act = F.gated_scores(X, Wcat, bcat, ww, bw, save=True)
    M, _, _, _ = F.segment_softmax_pool(X, s, offt)
    return F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM, (Wv, bv, Wu, bu)


def _close(name, a, b, n_rows):
    """Same arithmetic, different summation order (four column quarters per row in the fused kernel).  ds is a
    cancellation (g_i - dM.M with |g_i| ~ 30), the bias/score-weight gradients are sums of ds terms that cancel to ~0
    per bag: an absolute floor proportional to fp32 rounding of those magnitudes goes with the relative bound."""
    a = a.double().cpu().numpy().reshape(-1)
    b = b.double().cpu().numpy().reshape(-1)
    tol = 2e-4 * max(np.abs(b).max(), 1e-30) + (1e-4 if name == "dscores" else 2e-5 * n_rows)
    assert np.abs(a - b).max() <= tol, (name, float(np.abs(a - b).max()), tol)


def _two_kernel(F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM):
    ds, _ = F.segment_softmax_pool_bwd(X, s, offt, dM, M, want_attn=False)
    _, dWcat, dbcat, dww, dbw = F.gated_scores_bwd(X, Wcat, bcat, ww, bw, ds, None, dM, offt, False, gate_act=act)
    return ds, dWcat, dbcat, dww, dbw


CASES = [
    [1],                                   # one instance: a single, mostly empty k-block
    [31, 1, 33, 2, 64, 5],                 # boundaries inside k-blocks, 1-row bags
    [100, 257, 3000, 17, 1, 1, 999],       # several k-blocks per split
    [20000, 100, 7777],                    # long bags
    [5] * 200,                             # many bags per k-block (6 bag changes inside every block)
    [32 * 24 * 6],                         # exactly one block per (split, CTA)
    [32 * 24 * 6 + 1],
]


@pytest.mark.parametrize("lens", CASES)
def test_fused_backward_equals_two_kernel_backward(lens):
    F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM, _ = _setup(lens, 7)
    fused = F.gated_pool_bwd(X, s, offt, dM, M, ww, act)
    assert fused is not None, "milb200_gated_pool_bwd must cover bf16, D=192, L=1024"
    ref = _two_kernel(F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM)
    torch.cuda.synchronize()
    names = ("dscores", "dWcat", "dbcat", "dww", "dbw")
    for name, a, b in zip(names, fused, ref):
        _close(name, a, b, int(sum(lens)))


def test_fused_backward_is_repeatable():
    F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM, _ = _setup([4000, 123, 9000, 1500, 64, 31, 7000], 11)
    a = [t.clone() for t in F.gated_pool_bwd(X, s, offt, dM, M, ww, act)]
    for _ in range(5):
        b = F.gated_pool_bwd(X, s, offt, dM, M, ww, act)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_fused_backward_other_widths():
    for Lf in (512, 768):
        F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM, _ = _setup([300, 45, 1000, 2], 3, Lf=Lf)
        fused = F.gated_pool_bwd(X, s, offt, dM, M, ww, act)
        assert fused is not None
        ref = _two_kernel(F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM)
        for name, a, b in zip(("dscores", "dWcat", "dbcat", "dww", "dbw"), fused, ref):
            _close(name, a, b, 1347)


def test_module_backward_takes_the_fused_path_and_matches_oracle(monkeypatch):
    """nn.Module API (parameters require grad, the instances do not) with MILB200_FUSED_BWD=1: autograd must route through
    the fused kernel and agree with the float64 oracle's gradients within the bf16 bound (<= 1e-2)."""
    import mil_b200
    from mil_b200 import functional as F
    monkeypatch.setenv("MILB200_FUSED_BWD", "1")
    calls = []
    real = F.gated_pool_bwd
    monkeypatch.setattr(F, "gated_pool_bwd", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    lens = [700, 33, 1500, 64]
    off = mo.offsets_from_lengths(np.asarray(lens))
    p = mo.procedural_state(mo.abmil_shapes(1024), 99)
    m = mil_b200.ABMIL(None, L=1024).cuda().eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    gen = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn(int(off[-1]), 1024, device="cuda", generator=gen).bfloat16()
    offt = torch.from_numpy(off).cuda()
    dM = torch.randn(len(lens), 1024, device="cuda", generator=gen)
    M = m.forward_csr(X, offt)
    (M.float() * dM).sum().backward()
    torch.cuda.synchronize()
    assert calls, "autograd did not take the fused backward"
    pq = {k: (torch.from_numpy(v).bfloat16().float().numpy() if k.endswith("0.weight") else v) for k, v in p.items()}
    g = mo.abmil_backward_csr(pq, X.float().cpu().numpy(), off, dM.double().cpu().numpy(), need_dx=False)
    for k, prm in m.state_dict(keep_vars=True).items():
        assert prm.grad is not None, k
        got, want = prm.grad.double().cpu().numpy().reshape(-1), np.asarray(g[k]).reshape(-1)
        if k == "attention_weights.bias":          # = sum of ds = 0 analytically: only an absolute bound makes sense
            assert abs(got[0] - want[0]) <= 1e-4
        else:
            assert rel_err(got, want) <= 1e-2, k


def test_fused_dscores_match_oracle():
    lens = [257, 31, 1, 900]
    F, X, offt, Wcat, bcat, ww, bw, s, act, M, dM, (Wv, bv, Wu, bu) = _setup(lens, 21)
    ds = F.gated_pool_bwd(X, s, offt, dM, M, ww, act)[0]
    p = {"attention_V.0.weight": Wv.bfloat16().float().cpu().numpy(), "attention_V.0.bias": bv.cpu().numpy(),
         "attention_U.0.weight": Wu.bfloat16().float().cpu().numpy(), "attention_U.0.bias": bu.cpu().numpy(),
         "attention_weights.weight": ww.view(1, -1).cpu().numpy(), "attention_weights.bias": bw.cpu().numpy()}
    g = mo.abmil_backward_csr(p, X.float().cpu().numpy(), offt.cpu().numpy(), dM.double().cpu().numpy(), need_dx=False)
    assert rel_err(ds.cpu().numpy(), g["ds"]) <= 1e-2


def test_trainer_step_with_fused_backward_matches_the_default_step(monkeypatch):
    """AbmilTrainer (flat gradient buffer, fused Adam): one step with MILB200_FUSED_BWD=1 lands on the same parameters as
    the default two-kernel step."""
    from mil_b200.dp import AbmilTrainer
    lens = [900, 120, 3000, 31, 2048]
    off = torch.from_numpy(mo.offsets_from_lengths(np.asarray(lens))).cuda()
    gen = torch.Generator(device="cuda").manual_seed(8)
    X = torch.randn(int(off[-1]), 1024, device="cuda", generator=gen).bfloat16()
    out = []
    for flag in ("0", "1"):
        monkeypatch.setenv("MILB200_FUSED_BWD", flag)
        torch.manual_seed(3)
        tr = AbmilTrainer(L_feat=1024, D=192, lr=1e-3)
        tr.params.copy_(torch.randn(tr.numel, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4)) * 0.03)
        tr.step(X, off)
        g = tr.grads.clone()
        out.append((tr.params.clone(), g))
    (p0, g0), (p1, g1) = out
    scale = float(g0.abs().max())
    assert float((g0 - g1).abs().max()) <= 2e-4 * scale + 2e-5 * X.shape[0]
    assert float((p0 - p1).abs().max()) <= 2.1e-3          # Adam moves every weight by at most lr; sign flips of ~0 gradients


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_empty_bags_have_defined_outputs_and_zero_gradient_share(dtype):
    """Zero-length bags at the start, in the middle (also exactly on a slab boundary of the persistent pool kernel) and at
    the end of a packed batch: pooled vector 0, argmax -1, lse -inf; the other bags and every gradient are what the batch
    without the empty bags gives (both backward routes)."""
    from mil_b200 import functional as F
    lens_full = [0, 0, 700, 0, 4096, 0, 33, 0]
    lens = [l for l in lens_full if l > 0]
    off_full = torch.from_numpy(mo.offsets_from_lengths(np.asarray(lens_full))).cuda()
    off = torch.from_numpy(mo.offsets_from_lengths(np.asarray(lens))).cuda()
    gen = torch.Generator(device="cuda").manual_seed(13)
    X = torch.randn(sum(lens), L_FEAT, device="cuda", generator=gen).to(dtype)
    Wv = torch.randn(D, L_FEAT, device="cuda", generator=gen) * 0.03
    Wu = torch.randn(D, L_FEAT, device="cuda", generator=gen) * 0.03
    bv = torch.zeros(D, device="cuda")
    ww = torch.randn(D, device="cuda", generator=gen) * 0.3
    bw = torch.zeros(1, device="cuda")
    Wcat, bcat = F.pack_gate_weights(Wv, bv, Wu, bv, dtype)
    s, act = F.gated_scores(X, Wcat, bcat, ww, bw, save=True)
    M, _, am, lse = F.segment_softmax_pool(X, s, off)
    Mf, _, amf, lsef = F.segment_softmax_pool(X, s, off_full)
    keep = torch.tensor([i for i, l in enumerate(lens_full) if l > 0], device="cuda")
    gone = torch.tensor([i for i, l in enumerate(lens_full) if l == 0], device="cuda")
    assert torch.equal(Mf[keep], M) and torch.equal(amf[keep], am) and torch.equal(lsef[keep], lse)
    assert float(Mf[gone].abs().max()) == 0.0 and bool((amf[gone] == -1).all()) and bool(torch.isinf(lsef[gone]).all())
    dMf = torch.randn(len(lens_full), L_FEAT, device="cuda", generator=gen)
    dM = dMf[keep].contiguous()
    ds, _ = F.segment_softmax_pool_bwd(X, s, off, dM, M, want_attn=False)
    dsf, _ = F.segment_softmax_pool_bwd(X, s, off_full, dMf, Mf, want_attn=False)
    assert torch.equal(ds, dsf)
    if dtype == torch.bfloat16:
        a = F.gated_pool_bwd(X, s, off, dM, M, ww, act)
        b = F.gated_pool_bwd(X, s, off_full, dMf, Mf, ww, act)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
