"""model/dim1 of the reference: only the gated-attention MIL pools are on the hot path.
``gatedAttention`` is the name model/aggregator_wMask.py:24 imports; upstream never defines it —
it is the same gated pool with the ABMIL defaults (L=768, D=192, K=1)."""
from ...abmil import ABMIL, ABMIL_v2

gatedAttention = ABMIL

__all__ = ["ABMIL", "ABMIL_v2", "gatedAttention"]
