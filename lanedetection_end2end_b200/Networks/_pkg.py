"""Resolve the parent package whether ``Networks`` was imported as
``lanedetection_end2end_b200.Networks`` or as a top-level ``Networks`` (the way the
reference's main.py imports it, with this package directory first on sys.path)."""
import importlib
import os
import sys

_PKG = "lanedetection_end2end_b200"


def package():
    try:
        return importlib.import_module(_PKG)
    except ImportError:
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        if root not in sys.path:
            sys.path.insert(0, root)
        return importlib.import_module(_PKG)


def submodule(name):
    package()
    return importlib.import_module(_PKG + "." + name)


def capi():
    return submodule("_capi")
