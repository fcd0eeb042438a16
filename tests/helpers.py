"""Shared helpers for the test-suite (fixture loading, procedural inputs, tolerances)."""
import os

import numpy as np

from oracle.mil_oracle import procedural_state

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def rnd(seed, *shape, scale=1.0):
    """Same generator as tests/golden/make_golden.py:rnd."""
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


def digest(g):
    f = np.asarray(g, dtype=np.float64).reshape(-1)
    head = np.zeros(16)
    head[:min(16, f.size)] = f[:16]
    return np.concatenate([[f.sum(), (f * f).sum()], head])


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / den)


def check_grads(fx, grads, tol, prefix_filter=None):
    """Compare a {name: array} dict of gradients against the 'g:'/'d:' entries of a fixture.
    Digest entries compare sum (abs tol scaled by sqrt(sumsq)), sumsq (rel) and the 16-value head."""
    checked = 0
    for key, ref in fx.items():
        if key[:2] not in ("g:", "d:"):
            continue
        name = key[2:]
        if prefix_filter and not prefix_filter(name):
            continue
        if name.endswith("k_proj.bias") or name.endswith("attention_weights.bias"):
            # exactly-zero true gradient (softmax shift invariance): the reference holds float noise
            g = grads.get(name)
            if g is not None and ref.size and key[0] == "g":
                assert np.abs(np.asarray(g, dtype=np.float64)).max() <= max(1e-6, 10 * np.abs(ref).max())
            continue
        if ref.size == 0:
            g = grads.get(name)
            assert g is None or float(np.abs(np.asarray(g)).max()) == 0.0, f"{name}: expected no/zero grad"
            continue
        assert name in grads and grads[name] is not None, f"missing grad {name}"
        g = np.asarray(grads[name], dtype=np.float64)
        if key[0] == "g":
            scale = max(np.abs(ref).max(), 1e-30)
            err = np.abs(g - ref).max() / scale
            assert err <= tol, f"{name}: rel err {err:.3e} > {tol}"
        else:
            d = digest(g)
            rms = np.sqrt(max(ref[1], 1e-300))
            assert abs(d[0] - ref[0]) <= tol * rms * np.sqrt(g.size) + 1e-300, f"{name}: sum mismatch"
            assert abs(d[1] - ref[1]) <= 4 * tol * ref[1] + 1e-300, f"{name}: sumsq mismatch"
            hs = max(np.abs(ref[2:]).max(), rms / np.sqrt(g.size), 1e-30)
            assert np.abs(d[2:] - ref[2:]).max() / hs <= tol * 10, f"{name}: head mismatch"
        checked += 1
    return checked
