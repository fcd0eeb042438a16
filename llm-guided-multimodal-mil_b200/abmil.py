"""Gated-attention MIL pooling modules — drop-ins for model/dim1/ABMIL.py and ABMIL_v2.py.

Same constructor arguments, forward signature, outputs and ``state_dict`` keys as the reference
(``attention_V.0.*``, ``attention_U.0.*``, ``attention_weights.*``); the arithmetic runs in libmilb200's
sm_100a kernels: one fused tcgen05 GEMM + gate epilogue for the scores and one bandwidth-bound
segmented softmax/weighted-sum kernel over CSR bag offsets.  ``forward_csr`` is the B200-native entry:
a whole packed batch of ragged bags in one launch pair, equal to looping the reference over bags with
batch size 1 (train_ddp.py:75, test_ddp.py:73).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as F


class ABMIL(nn.Module):
    """model/dim1/ABMIL.py:6-64.  ``args`` is accepted and ignored exactly like upstream."""

    def __init__(self, args=None, L=768, D=192, K=1):
        super().__init__()
        self.L, self.D, self.K = L, D, K
        if K != 1:
            raise NotImplementedError("ABMIL: only K=1 attention branch exists upstream (ABMIL.py:7) and here")
        # nn.Linear modules are parameter containers only (state_dict ABI); their forward is never called
        self.attention_V = nn.Sequential(nn.Linear(self.L, self.D), nn.Tanh())
        self.attention_U = nn.Sequential(nn.Linear(self.L, self.D), nn.Sigmoid())
        self.attention_weights = nn.Linear(self.D, self.K)
        self.dropout1 = nn.Dropout(0.5)
        self.dropout2 = nn.Dropout(0.5)
        self.last_argmax = None   # int32 [B]: index (within each bag) of the largest attention score
        self.last_scores = None   # fp32 [total_n]: raw attention scores before the softmax

    def _params(self):
        return (self.attention_V[0].weight, self.attention_V[0].bias, self.attention_U[0].weight,
                self.attention_U[0].bias, self.attention_weights.weight, self.attention_weights.bias)

    def forward_csr(self, X, offsets, out_fp32=False):
        """X [total_n, L] packed instances, offsets int32 [B+1] on the device -> M [B, L] (X.dtype; fp32 with out_fp32)."""
        if X.dim() != 2 or X.shape[1] != self.L:
            raise L.MilB200Error(f"ABMIL.forward_csr: expected [total_n, {self.L}], got {tuple(X.shape)}")
        if self.training and self.dropout1.p > 0:
            X = F.dropout(X, self.dropout1.p)                                  # ABMIL.py:49
        M, am, s = F.abmil_pool_csr(X, offsets, *self._params(), out_fp32=out_fp32)
        self.last_argmax, self.last_scores = am, s
        return M

    def forward(self, x):
        x = x.squeeze(0)                                                       # ABMIL.py:48
        if x.dim() == 2:
            # CSR offsets [0, n] built on the device (a host list would cost a synchronous pageable H2D copy per bag)
            off = torch.arange(0, 2, dtype=torch.int32, device=x.device) * x.shape[0]
            return self.forward_csr(x, off)                                    # (K, L) = (1, L)
        if x.dim() == 3:
            # Upstream quirk: with a dense batch B>1 the softmax runs over the size-1 K axis, so
            # every weight is exactly 1, M[b] = sum_i x[b, i] and the attention parameters get no gradient.
            if self.training and self.dropout1.p > 0:
                x = F.dropout(x, self.dropout1.p)
            return F.dense_sum_pool(x)                                         # (B, 1, L)
        raise L.MilB200Error(f"ABMIL.forward: unsupported input rank {x.dim()}")


class ABMIL_v2(ABMIL):
    """model/dim1/ABMIL_v2.py:6-69: L fixed at 768, result concatenated with the BpRc class column."""

    def __init__(self, args=None):
        super().__init__(args, L=768, D=192, K=1)

    def forward(self, x, BpRc_class):
        M = super().forward(x)
        return torch.cat([M, BpRc_class.to(M.dtype)], dim=1)                   # ABMIL_v2.py:66 -> (1, 769)
