"""The reference's training loop (BP/main.py:239-340) through the drop-in modules with a real optimizer: on a fixed synthetic
batch with straight-lane ground truth the back-projection loss must fall (every kernel's forward AND backward, BatchNorm
running statistics, Dropout2d, Adam -- the functional check that the package trains, not just matches one step)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_loss_decreases_on_a_fixed_batch():
    import train_synthetic
    rep = train_synthetic.run(steps=40, batch=4)
    assert rep["skipped"] == 0
    assert rep["loss_last"] < 0.5 * rep["loss_first"], rep
    assert all(v == v and v < 1e12 for v in rep["losses"]), rep          # finite throughout
