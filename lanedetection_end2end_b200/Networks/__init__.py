"""Model registry -- the reference's plug-in seam (BP/Networks/__init__.py:8-19):
``define_model(mod='erfnet', **kwargs)`` selected by ``--mod``."""
from ._pkg import package as _package

_package()          # make `lanedetection_end2end_b200` importable when used as top-level `Networks`

from .ERFNet import Net  # noqa: E402

model_dict = {"erfnet": Net}


def allowed_models():
    return model_dict.keys()


def define_model(mod, **kwargs):
    if mod not in allowed_models():
        raise KeyError("The requested model: {} is not implemented".format(mod))
    return model_dict[mod](**kwargs)
