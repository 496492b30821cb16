"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Only works in the build container; /root/reference does not exist on the GPU
box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this.
Recipe = SURVEY.md Appendix A: the reference modules use top-level
``import Networks`` and need ``matplotlib`` at import time
(Backprojection_Loss/Networks/utils.py:17-21), which this image lacks -> stub it.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Backprojection_Loss", "Networks"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    p = types.ModuleType("matplotlib.pyplot")
    p.rcParams = {}
    m.pyplot = p
    sys.modules["matplotlib"] = m
    sys.modules["matplotlib.pyplot"] = p


def purge():
    for k in list(sys.modules):
        if k == "Networks" or k.startswith("Networks.") or k in ("Loss_crit",):
            del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.startswith(REFERENCE_ROOT)]


def import_reference(variant="Backprojection_Loss"):
    """Returns a namespace with the reference's Net, define_args, define_init_weights,
    backprojection_loss / Area_Loss, get_homography, Weighted_least_squares,
    ProjectiveGridGenerator, GELS for the given variant directory."""
    if not available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)
    _stub_matplotlib()
    purge()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, variant))
    ns = types.SimpleNamespace()
    import Networks  # noqa: F401
    from Networks import LSQ_layer, ERFNet, utils
    import Loss_crit
    ns.Networks = Networks
    ns.LSQ_layer = LSQ_layer
    ns.ERFNet = ERFNet
    ns.utils = utils
    ns.Loss_crit = Loss_crit
    if variant == "Backprojection_Loss":
        from Networks import gels
        ns.gels = gels
    return ns


def make_args(ns, extra=()):
    """argparse Namespace the reference's ``Net(args)`` consumes (utils.py:24-99)."""
    argv = ["--image_dir", "x", "--gt_dir", "y", "--no_cuda"] + list(extra)
    return ns.utils.define_args().parse_args(argv)
