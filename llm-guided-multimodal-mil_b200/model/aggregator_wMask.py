"""Late-fusion aggregator for mask-conditioned CT encoders — drop-in for model/aggregator_wMask.py:6-114
(``aggregator_wMask(args).forward(x_list, mask)``): concat[CT 768 | pathology pooled 768 | CI] -> Dropout,
Linear, ReLU, Dropout, Linear -> sigmoid.  Upstream imports a non-existent ``gatedAttention``;
here it is the gated pool with the ABMIL defaults.  ``forward_padded`` is the masked MIL entry of BASELINE
config 4: padded (B, Nmax, L) CT-slice bags with per-bag valid lengths, pooled without touching the padding."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F
from .._lib import MilB200Error
from .dim1 import gatedAttention
from .encoders import PrecomputedFeatures


class aggregator_wMask(nn.Module):
    def __init__(self, args, extractor_CT: nn.Module = None, extractor_CI: nn.Module = None):
        super().__init__()
        self.args = args
        mod = args.modality
        if "CT" in mod:                                                               # aggregator_wMask.py:12-20
            self.extractor_CT = extractor_CT if extractor_CT is not None else PrecomputedFeatures()
        if "pathology" in mod:                                                        # :22-28
            kind = getattr(args, "model_pathology", None)
            if kind == "ABMIL":
                self.extractor_pathology = gatedAttention(args)
            elif kind == "TransMIL":
                raise NotImplementedError("model_pathology='TransMIL' is outside the hot path (SURVEY F5)")
        if "CI" in mod:                                                               # :30-36
            self.extractor_CI = extractor_CI if extractor_CI is not None else PrecomputedFeatures()
        self.concat_feature_in = 0
        self.concat_feature_out = 0
        if "CT" in mod:
            self.concat_feature_in += 768
            self.concat_feature_out += 192
        if "pathology" in mod:
            self.concat_feature_in += 768
            self.concat_feature_out += 192
        if "CI" in mod:
            self.concat_feature_in += len(args.clinical_features)
        if ("CT" not in mod) and ("pathology" not in mod) and ("CI" in mod):          # :51-55
            self.fc = nn.Sequential(nn.Dropout(0.25), nn.Linear(self.concat_feature_in, args.num_classes))
        else:                                                                         # :66-70
            self.fc = nn.Sequential(nn.Dropout(0.25), nn.Linear(self.concat_feature_in, self.concat_feature_out), nn.ReLU(),
                                    nn.Dropout(0.25), nn.Linear(self.concat_feature_out, args.num_classes))

    def _ct(self, x, mask):
        if getattr(self.args, "model_CT", None) == "SwinUNETR_wMask":
            return self.extractor_CT(x, mask).squeeze(1)                              # :77
        return self.extractor_CT(torch.cat([x, mask], dim=1))                         # :79

    def _head(self, x):
        def drop(t, p):
            return F.dropout(t, p) if (self.training and p > 0) else t
        if len(self.fc) == 2:
            return F.linear(drop(x, self.fc[0].p), self.fc[1].weight, self.fc[1].bias, act="sigmoid")
        h = F.linear(drop(x, self.fc[0].p), self.fc[1].weight, self.fc[1].bias, act="relu")
        return F.linear(drop(h, self.fc[3].p), self.fc[4].weight, self.fc[4].bias, act="sigmoid")   # :114

    def forward(self, x_list, mask):
        mod = self.args.modality
        feats, i = [], 0
        if "CT" in mod:
            feats.append(self._ct(x_list[i], mask)); i += 1
        if "pathology" in mod:
            feats.append(self.extractor_pathology(x_list[i]).squeeze(1)); i += 1
        if "CI" in mod:
            feats.append(self.extractor_CI(x_list[i]).squeeze(1)); i += 1
        if not feats:
            raise MilB200Error(f"aggregator_wMask: unsupported modality {mod}")
        x = feats[0] if len(feats) == 1 else torch.cat(feats, dim=1)                  # :81,87,...
        return self._head(x)

    def forward_padded(self, pool, x_padded, lengths, other_feats=()):
        """Masked MIL over padded bags: x_padded (B, Nmax, L), lengths int32/int64 [B] valid rows per bag.  `pool`
        is a gated-attention module (e.g. the CT-slice pool).  Rows beyond lengths[b] never enter the softmax —
        the result equals the reference pool applied to each unpadded bag.  Returns the head output for
        cat([pooled, *other_feats], dim=1)."""
        B, Nmax, Lf = x_padded.shape
        lengths = lengths.to(device=x_padded.device, dtype=torch.int64)
        if int(lengths.min()) < 1 or int(lengths.max()) > Nmax:
            raise MilB200Error("forward_padded: lengths must lie in [1, Nmax]")
        keep = (torch.arange(Nmax, device=x_padded.device)[None, :] < lengths[:, None]).reshape(-1)
        rows = keep.nonzero(as_tuple=False).squeeze(1)
        packed = x_padded.reshape(B * Nmax, Lf).index_select(0, rows)                 # compaction = data movement only
        offsets = torch.zeros(B + 1, dtype=torch.int32, device=x_padded.device)
        offsets[1:] = lengths.cumsum(0).to(torch.int32)
        # the pooled vectors leave the kernel as fp32 and the (B x 1536 x 384) head runs in fp32: with bf16 bags the
        # instances' storage is then the only reduced-precision step (a bf16 head costs ~3e-2 on dL/dX for nothing)
        pooled = pool.forward_csr(packed, offsets, out_fp32=True)
        x = torch.cat([pooled, *[f.float() for f in other_feats]], dim=1) if other_feats else pooled
        return self._head(x)
