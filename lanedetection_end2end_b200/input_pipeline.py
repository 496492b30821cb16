"""Image side of the reference's loader on the GPU: what BP/Dataloader/Load_Data_new.py:127-131,
166-167,178-181 does per image with PIL on a DataLoader worker -- crop the bottom 640 rows, BILINEAR resize to
(resize, 2*resize), optional horizontal flip, ToTensor().float() -- as one launch over a batch of decoded frames.

    pre = FramePreprocessor(resize=256)                       # tables for 1280x720 TuSimple frames
    x = pre(frames_u8)                                        # [N,720,1280,3] uint8 cuda -> [N,3,256,512] float32
    x = pre(frames_u8, flip=mask, layout="nhwc4")             # or straight into the stem convolution's layout

The output is bit-identical to the reference's loader (Pillow's fixed-point resample is reproduced, csrc/input_pipe.cu).
JPEG decoding is outside this module (the frames arrive decoded); the label transforms of the loader (lane coordinates
/ 2.5, valid-point mask, horizon vector; :133-160) are O(100) scalars per image and stay on the host."""
import math

import numpy as np
import torch

from . import _capi

PRECISION_BITS = 32 - 8 - 2      # Pillow: 8 bits of pixel, 2 bits of headroom in a 32-bit accumulator


def resample_tables(in_size, out_size):
    """Per-axis tables of Pillow's ImagingResample for the BILINEAR (triangle, support 1) filter over [0, in_size):
    bounds int32 [out_size][2] (first source index, number of taps) and integer coefficients int32 [out_size][ksize]
    (Resample.c precompute_coeffs / normalize_coeffs_8bpc; float64 arithmetic in the same order as the C code)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / filterscale
    one = float(1 << PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [0.0] * xmax
        ww = 0.0
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * inv)
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(v * one - 0.5) if v < 0 else int(v * one + 0.5)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


class FramePreprocessor:
    def __init__(self, resize, frame_hw=(720, 1280), crop_rows=640, device="cuda"):
        self.resize, self.frame_hw, self.crop_rows = int(resize), tuple(frame_hw), int(crop_rows)
        H, W = self.frame_hw
        if self.crop_rows > H:
            raise ValueError("crop_rows %d exceeds the frame height %d" % (self.crop_rows, H))
        self.Ho, self.Wo = self.resize, 2 * self.resize
        xb, xk = resample_tables(W, self.Wo)
        yb, yk = resample_tables(self.crop_rows, self.Ho)
        dev = torch.device(device)
        self.tables = [torch.from_numpy(t).to(dev).contiguous() for t in (xb, xk, yb, yk)]
        self.kx, self.ky = xk.shape[1], yk.shape[1]

    def __call__(self, frames, flip=None, layout="nchw"):
        _capi.require_cuda(frames)
        H, W = self.frame_hw
        if frames.dtype != torch.uint8 or frames.dim() != 4 or tuple(frames.shape[1:]) != (H, W, 3):
            raise ValueError("expected uint8 frames [N,%d,%d,3], got %s %s" % (H, W, frames.dtype, tuple(frames.shape)))
        frames = frames.contiguous()
        N = frames.shape[0]
        if layout == "nchw":
            out, lay = torch.empty(N, 3, self.Ho, self.Wo, dtype=torch.float32, device=frames.device), 0
        elif layout == "nhwc4":
            out, lay = torch.empty(N, self.Ho, self.Wo, 4, dtype=torch.float32, device=frames.device), 1
        else:
            raise ValueError(layout)
        fl = None
        if flip is not None:
            fl = torch.as_tensor(flip, device=frames.device).to(torch.uint8).contiguous()
            if fl.numel() != N:
                raise ValueError("flip must have one entry per frame")
        xb, xk, yb, yk = self.tables
        _capi.call("lf_frame_preprocess", _capi.ptr(frames), N, H, W, H - self.crop_rows, self.crop_rows, _capi.ptr(xb), _capi.ptr(xk),
                   self.kx, _capi.ptr(yb), _capi.ptr(yk), self.ky, self.Ho, self.Wo, _capi.ptr(fl), lay, _capi.ptr(out),
                   _capi.stream_ptr(), nbytes=N * self.crop_rows * W * 3 + out.numel() * 4)
        return out
