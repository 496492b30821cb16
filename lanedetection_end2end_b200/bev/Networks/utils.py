"""``Networks.utils`` of the BEV variant: same helpers as the BP mirror."""
from ._pkg import bp

_u = bp("utils")
globals().update({k: getattr(_u, k) for k in dir(_u) if not k.startswith("__")})
