"""oracle/_ref (the reference's own ABMIL.py, copied unmodified by oracle/make_ref.py at build time) against the oracle's
restatements — a live pin in addition to the committed fixtures.  Skipped where the recipe has not run."""
import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo
from oracle import make_ref
from oracle import mil_oracle as mo


def test_reference_abmil_module_equals_both_oracles():
    Ref = make_ref.load_reference_abmil()
    if Ref is None:
        pytest.skip("oracle/_ref/ABMIL.py absent (python oracle/make_ref.py needs /root/reference)")
    L = 96
    p = mo.procedural_state(mo.abmil_shapes(L), 3)
    ref = Ref(None, L=L).eval()
    ref.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=True)
    x = torch.from_numpy(np.random.RandomState(4).standard_normal((1, 57, L)).astype(np.float32)).requires_grad_(True)
    M = ref(x)
    M.sum().backward()
    f = mo.abmil_forward(p, x.detach().numpy()[0])
    assert np.abs(M.detach().numpy() - f["M"]).max() <= 1e-5 * np.abs(f["M"]).max()
    sd = {"a." + k: torch.from_numpy(v).double().requires_grad_(True) for k, v in p.items()}
    xd = x.detach().double().requires_grad_(True)
    Mo = fo.abmil(sd, "a", xd)
    Mo.sum().backward()
    assert float((M.detach().double() - Mo.detach()).abs().max()) <= 1e-5 * float(Mo.abs().max())
    assert float((x.grad.double() - xd.grad).abs().max()) <= 1e-5 * float(xd.grad.abs().max())
    for k, v in ref.named_parameters():
        if k == "attention_weights.bias":
            continue
        g = sd["a." + k].grad
        assert float((v.grad.double() - g).abs().max()) <= 2e-5 * float(g.abs().max()), k
