"""MLPBlock — drop-in for model/sam/common.py:13-26 (Linear -> act -> Linear on the text tokens)."""
from __future__ import annotations

from typing import Type

import torch
import torch.nn as nn

from ... import functional as F
from ..._lib import MilB200Error

_ACT_NAMES = {nn.ReLU: "relu", nn.Tanh: "tanh", nn.Sigmoid: "sigmoid", nn.Identity: None}


class MLPBlock(nn.Module):
    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()

    def _act_name(self):
        kind = type(self.act)
        if kind not in _ACT_NAMES:
            # TwoWayTransformer always passes nn.ReLU (transformer.py:18,268); GELU is upstream's unused default
            raise MilB200Error(f"MLPBlock: activation {kind.__name__} has no fused epilogue in libmilb200 "
                               "(built: ReLU, Tanh, Sigmoid, Identity)")
        return _ACT_NAMES[kind]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = F.linear(x, self.lin1.weight, self.lin1.bias, act=self._act_name())     # common.py:26, act fused
        return F.linear(h, self.lin2.weight, self.lin2.bias)

    def emit(self, tape, x: int) -> int:
        return tape.linear(tape.linear(x, self.lin1, act=self._act_name()), self.lin2)
