"""Stand-ins for the reference's upstream encoders, which are OUT OF SCOPE (model/dim3/* CT CNNs,
the frozen CLIP text tower, simpleFCs).  The aggregators take the encoders' OUTPUTS as inputs: a CT feature
map (1, 512, c, h, w) / pooled CT features, and text embeddings (1, T, 512).  A caller that owns real
encoders passes them to the aggregator constructors (``extractor_CT=...``, ``clinic_extractor=...``)."""
import torch.nn as nn


class PrecomputedFeatures(nn.Module):
    """Pass-through: the 'input' already is the encoder's output.  Accepts the optional tumour mask argument of
    the *_wMask encoders (model/aggregator_wMask.py:77) and ignores it."""

    def forward(self, x, mask=None):
        return x
