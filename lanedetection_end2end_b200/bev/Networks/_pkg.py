import importlib
import os
import sys

_PKG = "lanedetection_end2end_b200"


def package():
    try:
        return importlib.import_module(_PKG)
    except ImportError:
        root = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
        if root not in sys.path:
            sys.path.insert(0, root)
        return importlib.import_module(_PKG)


def bp(name):
    """Module of the Backprojection-variant mirror (shared implementation)."""
    package()
    return importlib.import_module(_PKG + ".Networks." + name)
