"""model/utils_clip.py:6-8 — factory for the CLIP-joint-space aggregator."""


def get_model(args, **encoders):
    from .aggregator_clip import aggregator
    return aggregator(args, **encoders)
