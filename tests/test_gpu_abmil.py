"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI behind the
reference's nn.Module interface, against (a) the golden fixtures produced by the unmodified reference and
(b) the float64 oracle on identical seeded inputs.  Tolerances (BASELINE.json north_star): bag offsets and
attention argmax bit-exact; fp32 <= 1e-5 relative; bf16 <= 1e-2 relative (vs the oracle evaluated on the
same bf16-quantised operands)."""
import numpy as np
import pytest
import torch

from oracle import mil_oracle as mo
from tests.helpers import check_grads, load_golden, rel_err, rnd

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 1e-2


def _module(L, p, dtype=torch.float32):
    import mil_b200
    m = mil_b200.ABMIL(None, L=L).cuda().eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return m


def _grads(m):
    return {k: (v.grad.detach().float().cpu().numpy() if v.grad is not None else None)
            for k, v in m.state_dict(keep_vars=True).items()}


@pytest.mark.parametrize("name", ["abmil_L96_N37", "abmil_L768_N100", "abmil_L1024_N257", "abmil_L512_N300",
                                  "abmil_L1024_N1"])
def test_abmil_fp32_vs_reference_golden(name):
    """Dense (1,N,L) call exactly as the reference is called; compares with what the reference produced."""
    fx = load_golden(name)
    L, N, seed = int(fx["L"]), int(fx["N"]), int(fx["seed"])
    p = mo.procedural_state(mo.abmil_shapes(L), seed)
    m = _module(L, p)
    x = torch.from_numpy(rnd(seed + 100, 1, N, L)).cuda().requires_grad_(True)
    dM = torch.from_numpy(rnd(seed + 200, 1, L)).cuda()
    M = m(x)
    assert tuple(M.shape) == (1, L)
    (M * dM).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(M.detach().cpu().numpy(), fx["M"]) <= TOL_F32
    assert rel_err(m.last_scores.detach().cpu().numpy(), fx["s"]) <= TOL_F32
    assert int(m.last_argmax.item()) == int(fx["argmax"])                 # bit-exact index
    g = mo.abmil_backward(p, x.detach().cpu().numpy()[0], dM.detach().cpu().numpy())
    assert rel_err(x.grad.detach().cpu().numpy()[0], g["x"]) <= TOL_F32
    if N == 1:
        # a one-instance bag has softmax weight exactly 1: every attention gradient is exactly 0 in exact
        # arithmetic (the reference stores 0); the kernels may leave float noise from g_i - dM.M
        for k, g in _grads(m).items():
            assert float(np.abs(g).max()) <= 1e-5, k
    else:
        n = check_grads(fx, _grads(m), TOL_F32, prefix_filter=lambda k: k != "attention_weights.bias")
        assert n == 5
    assert abs(float(m.attention_weights.bias.grad)) <= 1e-5              # true gradient is exactly 0


@pytest.mark.parametrize("dtype,L,tol", [(torch.float32, 1024, TOL_F32), (torch.bfloat16, 1024, TOL_BF16),
                                         (torch.bfloat16, 768, TOL_BF16), (torch.bfloat16, 512, TOL_BF16),
                                         (torch.float32, 768, TOL_F32)])
def test_abmil_csr_ragged_vs_oracle(dtype, L, tol):
    """Packed ragged batch == looping the reference over bags with batch 1; fwd + all gradients."""
    p = mo.procedural_state(mo.abmil_shapes(L), 5)
    lens = np.concatenate([mo.ragged_lengths(9, 1, 900, 17), [1, 128, 129, 2500]])
    off = mo.offsets_from_lengths(lens)
    X = rnd(23, int(off[-1]), L)
    m = _module(L, p)
    Xt = torch.from_numpy(X).cuda().to(dtype).requires_grad_(True)
    offt = torch.from_numpy(off).cuda()
    M = m.forward_csr(Xt, offt)
    dM = rnd(29, len(lens), L)
    (M.float() * torch.from_numpy(dM).cuda()).sum().backward()
    torch.cuda.synchronize()
    # oracle on the operands the kernels actually saw
    pq = {k: (torch.from_numpy(v).to(dtype).float().numpy() if k.endswith("0.weight") else v) for k, v in p.items()}
    Xq = Xt.detach().float().cpu().numpy()
    Mr, sr, amr = mo.abmil_forward_csr(pq, Xq, off)
    assert mo.score_margin(sr, off) > 1e-4, "seed gives a near-tie; argmax would not be well-posed"
    gr = mo.abmil_backward_csr(pq, Xq, off, dM)
    assert np.array_equal(offt.detach().cpu().numpy(), off)                         # offsets untouched, bit-exact
    assert m.last_argmax.detach().cpu().numpy().tolist() == amr.tolist()            # bit-exact argmax
    assert rel_err(M.detach().float().cpu().numpy(), Mr) <= tol
    assert rel_err(m.last_scores.detach().cpu().numpy(), sr) <= tol
    assert rel_err(Xt.grad.float().cpu().numpy(), gr["x"]) <= tol
    g = _grads(m)
    for k in p:
        if k == "attention_weights.bias":
            assert abs(float(g[k][0])) <= 1e-4
            continue
        assert rel_err(g[k], gr[k]) <= tol, k


def test_abmil_dense_batch_quirk_matches_reference():
    """dense (B>1,N,L) input -> plain sum pool, zero gradient to the attention parameters."""
    fx = load_golden("abmil_dense_batched_B3")
    p = mo.procedural_state(mo.abmil_shapes(96), int(fx["seed"]))
    m = _module(96, p)
    x = torch.from_numpy(rnd(121, 3, 17, 96)).cuda().requires_grad_(True)
    M = m(x)
    assert tuple(M.shape) == (3, 1, 96)
    assert rel_err(M.detach().cpu().numpy(), fx["M"]) <= TOL_F32
    M.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    assert all(v.grad is None for v in m.parameters())


def test_abmil_v2_matches_reference_golden():
    import mil_b200
    fx = load_golden("abmil_v2_N29")
    seed, N = int(fx["seed"]), int(fx["N"])
    p = mo.procedural_state(mo.abmil_shapes(768), seed)
    m = mil_b200.ABMIL_v2(None).cuda().eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    x = torch.from_numpy(rnd(seed + 100, 1, N, 768)).cuda()
    out = m(x, torch.tensor([[1.0]]).cuda())
    assert tuple(out.shape) == (1, 769)
    assert rel_err(out.detach().cpu().numpy(), fx["M"]) <= TOL_F32


def test_masked_padded_bags_equal_unpadded_reference():
    """cfg 4: padded (B,Nmax,L) CT-slice bags with valid lengths == reference ABMIL on each unpadded bag."""
    L, B, Nmax = 768, 8, 160
    p = mo.procedural_state(mo.abmil_shapes(L), 41)
    m = _module(L, p)
    Xpad = rnd(43, B, Nmax, L)
    lens = mo.ragged_lengths(B, 40, 160, 44)
    Mr = mo.abmil_forward_masked(p, Xpad, lens)
    packed = np.concatenate([Xpad[b, :int(n)] for b, n in enumerate(lens)])
    M = m.forward_csr(torch.from_numpy(packed).cuda(), torch.from_numpy(mo.offsets_from_lengths(lens)).cuda())
    assert rel_err(M.detach().cpu().numpy(), Mr) <= TOL_F32


def test_train_mode_dropout_is_statistical_and_regenerated():
    """train-mode dropout(0.5) on the bag; mask cannot bit-match torch, so check its statistics, the
    1/(1-p) scaling and that backward regenerates the same mask."""
    from mil_b200 import functional as F
    x = torch.ones(4096, 512, device="cuda", requires_grad=True)
    torch.manual_seed(7)
    y = F.dropout(x, 0.5)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.5) < 0.01
    assert torch.all((y == 0) | (y == 2.0))
    y.sum().backward()
    assert torch.equal(x.grad, y.detach())                                # same mask, same scale
    torch.manual_seed(7)
    assert torch.equal(F.dropout(x, 0.5).detach(), y.detach())            # seeded => reproducible
    assert not torch.equal(F.dropout(x, 0.5).detach(), y.detach())


def test_tensor_core_and_ffma_paths_agree():
    """The tcgen05 bf16 kernels against the FFMA kernels on the same bf16 operands (independent code paths)."""
    import subprocess, sys, os
    code = r"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import mil_b200
from oracle import mil_oracle as mo
p = mo.procedural_state(mo.abmil_shapes(1024), 3)
m = mil_b200.ABMIL(None, L=1024).cuda().eval()
m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
lens = mo.ragged_lengths(7, 50, 3000, 9); off = mo.offsets_from_lengths(lens)
X = torch.from_numpy(np.random.RandomState(1).standard_normal((int(off[-1]), 1024)).astype(np.float32)).cuda().bfloat16()
X.requires_grad_(True)
M = m.forward_csr(X, torch.from_numpy(off).cuda()); M.float().sum().backward(); torch.cuda.synchronize()
np.savez(sys.argv[1], M=M.detach().float().cpu().numpy(), s=m.last_scores.detach().cpu().numpy(), dX=X.grad.float().cpu().numpy(),
         dW=m.attention_V[0].weight.grad.detach().cpu().numpy(), am=m.last_argmax.detach().cpu().numpy())
"""
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for force in ("0", "1"):
        f = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, MILB200_FORCE_SIMT=force)
        subprocess.run([sys.executable, "-c", code, f], check=True, cwd=root, env=env, timeout=300)
        outs.append(dict(np.load(f)))
        os.remove(f)
    a, b = outs
    assert np.array_equal(a["am"], b["am"])
    for k in ("M", "s", "dX", "dW"):
        assert rel_err(a[k], b[k]) <= TOL_BF16, k
