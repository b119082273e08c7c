from estimator.registry import MODELS


def build_model(cfg):
    """`estimator/models/builder.py:6-8`"""
    return MODELS.build(cfg)
