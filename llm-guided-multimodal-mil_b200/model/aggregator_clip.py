"""CLIP-joint-space aggregator — drop-in for model/aggregator_clip.py:6-118 (``aggregator(args).forward(x_list)``):
gated-attention MIL pool over the pathology bag -> ``fc_pathology`` (Dropout .25, Linear 768->512, ReLU), CT feature
-> ``fc_CT``, mean of the two -> head -> sigmoid; the 512-d features are returned for the external CLIP loss
(``mil_b200.clip_loss``).  ``forward_csr`` is the B200-native batched entry: B ragged bags in one launch set.
The CT encoder is upstream of the hot path: ``x_list[0]`` is its OUTPUT (B, 512) unless one is injected."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F
from .._lib import MilB200Error
from ..abmil import ABMIL, ABMIL_v2
from .encoders import PrecomputedFeatures


class aggregator(nn.Module):
    def __init__(self, args, extractor_CT: nn.Module = None):
        super().__init__()
        self.args = args
        self.concat_feature_in = 0
        self.concat_feature_out = 0
        if "CT" in args.modality:                                                     # aggregator_clip.py:14-32
            self.concat_feature_in_CT = 512
            self.concat_feature_mid_CT = 512
            self.concat_feature_out += self.concat_feature_mid_CT
            self.extractor_CT = extractor_CT if extractor_CT is not None else PrecomputedFeatures()
        if "pathology" in args.modality:                                              # :34-53
            kind = getattr(args, "model_pathology", None)
            if kind == "ABMIL":
                self.concat_feature_in_pathology = 768
                self.extractor_pathology = ABMIL(args)
            elif kind == "ABMIL_v2":
                self.concat_feature_in_pathology = 768 + 1
                self.extractor_pathology = ABMIL_v2(args)
            elif kind == "TransMIL":
                raise NotImplementedError("model_pathology='TransMIL' is outside the hot path (SURVEY F5)")
            else:
                raise MilB200Error(f"aggregator_clip: unknown model_pathology {kind!r}")
            self.concat_feature_mid_pathology = 512
            self.concat_feature_out += self.concat_feature_mid_pathology
        if len(args.modality) == 1:                                                   # :57-61
            self.concat_feature_mid = (self.concat_feature_in_CT if "CT" in args.modality
                                       else self.concat_feature_in_pathology)
        elif len(args.modality) == 2:                                                 # :62-71
            if "CT" in args.modality:
                self.fc_CT = nn.Sequential(nn.Dropout(0.25), nn.Linear(self.concat_feature_in_CT, self.concat_feature_mid_CT),
                                           nn.ReLU())
            if "pathology" in args.modality:
                self.fc_pathology = nn.Sequential(nn.Dropout(0.25),
                                                  nn.Linear(self.concat_feature_in_pathology, self.concat_feature_mid_pathology),
                                                  nn.ReLU())
            self.concat_feature_mid = self.concat_feature_mid_CT
        self.fc = nn.Sequential(nn.Dropout(0.25), nn.Linear(self.concat_feature_mid, args.num_classes))   # :72-75

    def _drop_linear(self, seq, x, act):
        if self.training and seq[0].p > 0:
            x = F.dropout(x, seq[0].p)
        return F.linear(x, seq[1].weight, seq[1].bias, act=act)

    def _ct_features(self, x):
        kind = getattr(self.args, "model_CT", None)
        if kind == "SwinUNETR":
            return self.extractor_CT(x).squeeze(1)                                    # :83-84
        if kind == "MViT":
            return self.extractor_CT(x.squeeze(1))                                    # :85-86
        return self.extractor_CT(x)

    def _mix_and_head(self, x_CT, x_pathology):
        x_CT = self._drop_linear(self.fc_CT, x_CT, "relu")                            # :89
        x_pathology = self._drop_linear(self.fc_pathology, x_pathology, "relu")       # :92
        if x_CT.shape != x_pathology.shape:
            raise MilB200Error(f"aggregator_clip: CT features {tuple(x_CT.shape)} vs pathology {tuple(x_pathology.shape)}")
        x = F.axpby(x_CT, x_pathology, 0.5, 0.5)                                      # :94 (x_CT + x_pathology) / 2
        return x_CT, x_pathology, self._drop_linear(self.fc, x, "sigmoid")            # :96

    def forward(self, x_list):
        mod = self.args.modality
        if "CT" in mod and "pathology" in mod:
            x_CT = self._ct_features(x_list[0])
            x_pathology = self.extractor_pathology(x_list[1]).squeeze(1)              # :91
            return self._mix_and_head(x_CT, x_pathology)
        if "CT" in mod:
            x_CT = self._ct_features(x_list[0])
            return x_CT, self._drop_linear(self.fc, x_CT, "sigmoid")                  # :107
        if "pathology" in mod:
            if getattr(self.args, "model_pathology", None) == "ABMIL_v2":
                x_pathology = self.extractor_pathology(x_list[0], x_list[1]).squeeze(1)   # :112
            else:
                x_pathology = self.extractor_pathology(x_list[0]).squeeze(1)          # :115
            return x_pathology, self._drop_linear(self.fc, x_pathology, "sigmoid")    # :118
        raise MilB200Error(f"aggregator_clip: unsupported modality {mod}")

    def forward_csr(self, x_CT, X, offsets):
        """B bags at once: x_CT (B, 512) CT features, X [total_n, 768] packed instances, offsets int32 [B+1].
        Equals stacking ``forward([x_CT[b:b+1], X[offsets[b]:offsets[b+1]][None]])`` over b."""
        pooled = self.extractor_pathology.forward_csr(X, offsets)                     # (B, 768)
        if "CT" in self.args.modality:
            return self._mix_and_head(self._ct_features(x_CT), pooled)
        return pooled, self._drop_linear(self.fc, pooled, "sigmoid")
