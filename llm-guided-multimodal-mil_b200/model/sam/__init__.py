from .common import MLPBlock
from .transformer import Attention, TwoWayAttentionBlock, TwoWayTransformer

__all__ = ["MLPBlock", "Attention", "TwoWayAttentionBlock", "TwoWayTransformer"]
