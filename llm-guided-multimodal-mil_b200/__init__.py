"""mil_b200 — B200-native (sm_100a) implementation of the multimodal MIL aggregator hot path of
KyleKWKim/LLM-guided-Multimodal-MIL behind the reference's nn.Module interfaces.

The directory name carries hyphens, so import it through the repo-root shim: ``import mil_b200``.
"""
from . import _lib
from ._lib import MilB200Error, launch_count, lib
from . import functional
from .abmil import ABMIL, ABMIL_v2

__all__ = ["ABMIL", "ABMIL_v2", "functional", "lib", "launch_count", "MilB200Error"]
