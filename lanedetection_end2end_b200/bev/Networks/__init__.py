"""Model registry of the BEV variant (Birds_Eye_View_Loss/Networks/__init__.py:8-20)."""
from ._pkg import package as _package

_package()

from .ERFNet import Net  # noqa: E402

model_dict = {"erfnet": Net}


def allowed_models():
    return model_dict.keys()


def define_model(mod, **kwargs):
    if mod not in allowed_models():
        raise KeyError("The requested model: {} is not implemented".format(mod))
    return model_dict[mod](**kwargs)
