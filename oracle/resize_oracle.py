"""CPU restatement of the image side of the reference's loader (TEST INFRASTRUCTURE ONLY).

BP/Dataloader/Load_Data_new.py:127-131,166-167,178-181: crop the bottom 640 rows (F.crop), F.resize to (resize, 2*resize)
with PIL BILINEAR, optional F.hflip, ToTensor().float() (uint8 / 255).  The arithmetic lives in Pillow (a third-party
dependency of the reference, not under /root/reference; 12.2.0 in this image): ImagingResample in src/libImaging/Resample.c
-- a separable triangle filter whose support grows with the down-scale factor (antialiasing), evaluated per axis in
fixed point: coefficients (int)(w * 2^22 +- 0.5), accumulator 2^21 + sum(pixel * coeff), result clip8(acc >> 22);
horizontal pass first, 8-bit intermediate image, then the vertical pass.  Restated here in numpy (loops over output
columns / rows only); pinned bit-for-bit against Pillow itself by tests/test_input_pipeline_cpu.py."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the whole axis."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(xmax)
        ww = 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((len(bounds),) + img.shape[1:], dtype=np.uint8)
    for xx, (xmin, xmax) in enumerate(bounds):
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += img[xmin + x] * kk[xx, x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, out_h, out_w):
    """img uint8 [H,W,C] -> uint8 [out_h,out_w,C], Pillow's Image.resize((out_w, out_h), BILINEAR)."""
    H, W, _ = img.shape
    t = img
    if W != out_w:
        t = _pass(t, *_coeffs(W, out_w), axis=1)
    if H != out_h:
        t = _pass(t, *_coeffs(H, out_h), axis=0)
    return t


def preprocess(frame, resize, crop_rows=640, flip=False):
    """One decoded RGB frame uint8 [h,w,3] -> float32 [3, resize, 2*resize] exactly as the loader hands it to the model."""
    h = frame.shape[0]
    img = resize_bilinear_u8(frame[h - crop_rows:], resize, 2 * resize)
    if flip:
        img = img[:, ::-1]
    return np.ascontiguousarray(np.transpose(img.astype(np.float32) / np.float32(255.0), (2, 0, 1)))
