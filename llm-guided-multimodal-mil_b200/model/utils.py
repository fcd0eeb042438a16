"""model/utils.py:6-12 — the model factory used by train_ddp.py:68 / test_ddp.py:67."""


def get_model(args, **encoders):
    if "CT" in args.modality and "wMask" in getattr(args, "model_CT", ""):
        from .aggregator_wMask import aggregator_wMask
        return aggregator_wMask(args, **encoders)
    from .aggregator import aggregator
    return aggregator(args, **encoders)
