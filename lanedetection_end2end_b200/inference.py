"""Inference-side companions of the hot path: what the reference's ``test.py`` does with the
curve coefficients after ``model(input, ...)``.

  Projections                BP/test.py:132-186  sample every fitted curve at the 56 TuSimple rows in BEV space and
                                                 project the samples back into the image with M^-1 -> x coordinates
  lanes_from_predictions     BP/test.py:66-86    gate lanes by the line-type / horizon heads, clip to the image
  write_tusimple_predictions BP/test.py:88-99    one JSON line per image in the TuSimple submission format

``Projections.compute_coordinates`` keeps the reference's name, argument and result ([B, order+1, 1] float64 -> [B, 56]
float64, in 1280x720 pixels); on a CUDA tensor it is ONE launch of lf_backproj_loss (csrc/loss.cu, the kernel that
evaluates the training loss: same per-point code, x_cal output only) instead of two bmm's, a stack/permute and two
divisions; ``compute_all`` does all lanes of the batch in that one launch.
"""
import ctypes
import json

import numpy as np
import torch

if __package__:
    from . import _capi
    from .Loss_crit import backprojection_loss
else:                                   # imported top-level next to the reference's main.py / test.py
    import _capi
    from Loss_crit import backprojection_loss


def resize_coordinates(array):
    """256x512 network coordinates -> 1280-wide TuSimple frames (BP/test.py:20-21)."""
    return array * 2.5


class Projections:
    """Back-projected lane x-coordinates at the TuSimple h_samples (BP/test.py:132-186)."""

    def __init__(self, options):
        if options.order > 3:
            raise NotImplementedError("Requested order {} for polynomial fit is not implemented".format(options.order))
        self._crit = backprojection_loss(options)      # owns Y56, y', M^-1 (the same constants, BP/Loss_crit.py:166-188)
        self.M, self.M_inv = self._crit.M, self._crit.M_inv

    def compute_all(self, betas):
        """betas: sequence of L tensors [B, order+1, 1] (Net.forward's beta0..3, None entries skipped) -> [B, L, 56] f64."""
        betas = [b for b in betas if b is not None]
        beta = torch.stack([b.reshape(b.shape[0], -1) for b in betas], 1).double().contiguous()       # [B, L, n]
        B, L, n = beta.shape
        if not beta.is_cuda:
            raise _capi.LanefitError("Projections runs on CUDA tensors (no CPU fallback by design)")
        dev = beta.device
        ones = torch.ones(B, L, 56, dtype=torch.float64, device=dev)
        zeros = torch.zeros(B, L, 56, dtype=torch.float64, device=dev)
        lane = torch.empty(L, dtype=torch.float64, device=dev)
        loss = torch.empty(1, dtype=torch.float64, device=dev)
        xcal = torch.empty(B, L, 56, dtype=torch.float64, device=dev)
        host = self._crit._fused_host_constants()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _capi.call("lf_backproj_loss", host[0].ctypes.data, host[1].ctypes.data, host[2].ctypes.data, p(beta), p(zeros), p(ones),
                   B, L, n, p(lane), p(loss), None, p(xcal), p(self._crit._fused_ticket(dev)), _capi.stream_ptr())
        return resize_coordinates(xcal)

    def compute_coordinates(self, params):
        """params [B, order+1, 1] -> x coordinates [B, 56] in the 1280-wide frame (reference signature)."""
        return self.compute_all([params])[:, 0]


def lanes_from_predictions(x_cal, line_pred=None, horizon_pred=None):
    """x_cal [B, 4, 56] (Projections.compute_all) -> integer lane lists as the reference writes them (BP/test.py:66-86):
    lanes the line-type head switches off, samples above the predicted horizon and samples outside [0, 1279] become -2.
    line_pred [B, 4] in the head's order (rounded sigmoid); horizon_pred [B] in 720p rows (multiples of 10)."""
    lanes = x_cal.clone()
    if line_pred is not None:
        lp = line_pred[:, [1, 2, 0, 3]]
        lanes[(1 - lp[:, :, None]).bool().expand_as(lanes)] = -2
    if horizon_pred is not None:
        bounds = ((horizon_pred - 160) / 10)
        for k, bound in enumerate(bounds):
            lanes[k, :, :max(int(bound.item()), 0)] = -2
    lanes[lanes > 1279] = -2
    lanes[lanes < 0] = -2
    return np.int_(np.round(lanes.detach().cpu().numpy())).tolist()


def write_tusimple_predictions(json_file, gt_lines, lanes_pred, first_index, run_time=20):
    """Append one TuSimple submission line per image (BP/test.py:88-99; `run_time` is the reference's constant 20)."""
    for j, lanes in enumerate(lanes_pred):
        line = dict(gt_lines[first_index + j])
        line["lanes"] = lanes
        line["run_time"] = run_time
        json.dump(line, json_file)
        json_file.write("\n")
