"""Birds_Eye_View_Loss/Loss_crit.py mirror: Area_Loss (:78-134) and MSE_Loss (:137-150) are shared
with the BP mirror."""
import importlib
import os
import sys

try:
    _L = importlib.import_module("lanedetection_end2end_b200.Loss_crit")
except ImportError:
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
    _L = importlib.import_module("lanedetection_end2end_b200.Loss_crit")

Area_Loss = _L.Area_Loss
MSE_Loss = _L.MSE_Loss
define_loss_crit = _L.define_loss_crit
