"""Model registry -- the reference's plug-in seam (BP/Networks/__init__.py:8-19):
``define_model(mod='erfnet', **kwargs)`` selected by ``--mod``."""
from ._pkg import package as _package

_package()          # make `lanedetection_end2end_b200` importable when used as top-level `Networks`

from . import ERFNet as _erfnet  # noqa: E402

Net = _erfnet.Net
model_dict = {"erfnet": Net}     # --mod value -> constructor


def allowed_models():
    """Names accepted by ``--mod`` (view of the registry keys, like the reference returns)."""
    return model_dict.keys()


def define_model(mod, **kwargs):
    try:
        ctor = model_dict[mod]
    except KeyError:
        raise KeyError("The requested model: {} is not implemented".format(mod)) from None
    return ctor(**kwargs)
