"""Model registry of the BEV variant (Birds_Eye_View_Loss/Networks/__init__.py:8-20): the same seam as the
back-projection package, bound to the BEV flavour of the network (2-tuple ERFNet, normalised grid)."""
from ._pkg import package as _package

_package()

from . import ERFNet as _erfnet_bev  # noqa: E402

Net = _erfnet_bev.Net
model_dict = {"erfnet": Net}     # --mod value -> constructor


def allowed_models():
    """Names accepted by ``--mod``."""
    return model_dict.keys()


def define_model(mod, **kwargs):
    try:
        ctor = model_dict[mod]
    except KeyError:
        raise KeyError("The requested model: {} is not implemented".format(mod)) from None
    return ctor(**kwargs)
