"""Mirror of the reference's ``model`` package for the hot path: same module paths below ``mil_b200.model``
(``aggregator``, ``aggregator_clip``, ``aggregator_wMask``, ``utils.get_model``, ``utils_clip.get_model``,
``sam.transformer``, ``dim1``)."""
from . import aggregator, aggregator_clip, aggregator_wMask, dim1, sam, utils, utils_clip  # noqa: F401
from .utils import get_model  # noqa: F401
