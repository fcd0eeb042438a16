"""CPU tests: the C-ABI library builds/loads and exports every symbol include/milb200.h declares; the product
package has no CPU fallback and never imports the oracle."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "milb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(milb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    import mil_b200
    h = mil_b200.lib()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, f"declared in include/milb200.h but not exported: {missing}"
    assert h.milb200_version() == 200
    assert h.milb200_launch_count() >= 0


def test_ctypes_signatures_cover_the_header():
    import mil_b200
    from mil_b200 import _lib
    assert set(_declared()) <= set(_lib.SIGNATURES), sorted(set(_declared()) - set(_lib.SIGNATURES))


def test_no_cpu_fallback():
    """Feeding CPU tensors must raise, not silently compute."""
    import mil_b200
    m = mil_b200.ABMIL(None, L=64)
    with pytest.raises(mil_b200.MilB200Error):
        m(torch.randn(1, 5, 64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "llm-guided-multimodal-mil_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", ""), f"{f} mentions oracle"


def test_state_dict_abi_matches_reference_names():
    import mil_b200
    m = mil_b200.ABMIL(None, L=1024)
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "attention_V.0.weight": (192, 1024), "attention_V.0.bias": (192,),
        "attention_U.0.weight": (192, 1024), "attention_U.0.bias": (192,),
        "attention_weights.weight": (1, 192), "attention_weights.bias": (1,)}
    assert sum(p.numel() for p in m.parameters()) == 393793
