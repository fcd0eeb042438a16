"""CLIP-style image-text similarity logits and the losses around the aggregators.

  * ``CLIPLogits``        — the tail of clip/model.py:354-368 (``CLIP.forward``): L2-normalise both feature sets,
                            ``exp(logit_scale) * I @ T^T`` and its transpose; ``logit_scale`` is the same parameter
                            (init ln(1/0.07), clip/model.py:291).  The image/text towers themselves are frozen upstream
                            encoders (out of scope): features go in.
  * ``CLIPloss_v1``       — utils.py:247-284 given the frozen CLIP text features ``[b, I, 512]`` that upstream computes
                            with ``encode_text`` inside the loss: logits ``[I, b, b]`` (no normalisation, no temperature)
                            + cross-entropy over dim 1 against the identity.
  * ``bce_loss`` / ``cosine_embedding_loss`` — train_ddp.py:99,102,319-326.
All of them are single fused libmilb200 calls (forward and backward computed together where the op is tiny)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as F


class CLIPLogits(nn.Module):
    def __init__(self):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

    def forward(self, image_features, text_features):
        """-> (logits_per_image [b_i, b_t], logits_per_text [b_t, b_i]) fp32."""
        return F.clip_logits(image_features, text_features, self.logit_scale)


class CLIPloss_v1(nn.Module):
    """``forward(output [b,512], text_features [b, I, 512]) -> loss``; ``last_logits`` keeps the [I,b,b] logits."""

    def __init__(self, args=None):
        super().__init__()
        self.args = args
        self.clinical_info = getattr(args, "clinical_features", None)
        self.last_logits = None

    def forward(self, output, text_features):
        loss, logits = F.cliploss_v1(output, text_features)
        self.last_logits = logits
        return loss


def bce_loss(prob_logits, target):
    """sigmoid + nn.BCELoss(mean) fused; takes the PRE-sigmoid head output.  Returns (loss, prob)."""
    return F.sigmoid_bce(prob_logits, target)


def cosine_embedding_loss(a, b):
    """nn.CosineEmbeddingLoss()(a, b, ones) = mean(1 - cos(a_i, b_i))."""
    return F.cosine_embedding_loss(a, b)
