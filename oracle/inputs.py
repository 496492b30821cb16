"""Deterministic synthetic inputs shared by the golden generator, the tests and
bench.py (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Everything is generated with numpy's PCG64 ``default_rng(seed)`` so that the
build container (where the reference is importable and goldens are made) and
the GPU box (where only /root/repo exists) regenerate bit-identical arrays;
each golden file stores a sha256 of its inputs so drift is detected.

Workload definitions follow SURVEY.md section 8(d):
  * images: ``rand(B,3,H,W)`` float32 in [0,1) (the loader's range,
    Backprojection_Loss/Dataloader/Load_Data_new.py:184)
  * loss targets: ``x_gt = rand(B,56)*500`` f64, ``valid = ones`` with the
    first 8 columns zeroed (Load_Data_new.py:140-141)
  * LSQ stress maps: lane-like ridges ``exp(-((x-c-s(y-128))/6)^2/2) +
    0.02*U(0,1)`` (config 5)
"""
import hashlib
import math

import numpy as np


def sha256_of(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def make_images(B, H, W, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((B, 3, H, W), dtype=np.float32)


def make_loss_targets(B, nlanes=4, seed=1):
    """x_gt [B,nlanes,56] f64 in [0,500), valid [B,nlanes,56] f64 (first 8 cols 0)."""
    rng = np.random.default_rng(seed)
    x_gt = rng.random((B, nlanes, 56)) * 500.0
    valid = np.ones((B, nlanes, 56), dtype=np.float64)
    valid[:, :, :8] = 0.0
    return x_gt, valid


def make_lane_maps(B, L, H, W, seed=0, noise=0.02, dtype=np.float32):
    """Raw decoder-like maps o[B,L,H,W]: one slanted Gaussian ridge per map plus
    uniform noise.  The LSQ layer squares them (w = o**2)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(100.0 * W / 512.0, 400.0 * W / 512.0, size=(B, L, 1, 1))
    s = rng.uniform(-0.75, 0.75, size=(B, L, 1, 1))
    yy = np.arange(H, dtype=np.float64).reshape(1, 1, H, 1)
    xx = np.arange(W, dtype=np.float64).reshape(1, 1, 1, W)
    ridge = np.exp(-0.5 * ((xx - c - s * (yy - H / 2.0)) / 6.0) ** 2)
    o = ridge + noise * rng.random((B, L, H, W))
    return o.astype(dtype)


def make_uniform_maps(B, L, H, W, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((B, L, H, W), dtype=np.float32)


def make_grad_beta(B, L, order, seed=7):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, L, order + 1))


# --------------------------------------------------------------------------
# ERFNet parameters (names/shapes = the reference state_dict)
# --------------------------------------------------------------------------

def erfnet_param_shapes(in_channels=3, nclasses=2):
    """Ordered (name, shape, kind) list for ``net.*`` of the reference model
    (Backprojection_Loss/Networks/ERFNet.py:11-168).  kind in {conv, convT, bn_w,
    bn_b, bias}."""
    out = []

    def conv(prefix, co, ci, kh, kw):
        out.append((prefix + ".weight", (co, ci, kh, kw), "conv"))
        out.append((prefix + ".bias", (co,), "bias"))

    def convT(prefix, ci, co, kh, kw):
        out.append((prefix + ".weight", (ci, co, kh, kw), "convT"))
        out.append((prefix + ".bias", (co,), "bias"))

    def bn(prefix, c):
        out.append((prefix + ".weight", (c,), "bn_w"))
        out.append((prefix + ".bias", (c,), "bn_b"))

    def down(prefix, ci, co):
        conv(prefix + ".conv", co - ci, ci, 3, 3)
        bn(prefix + ".bn", co)

    def nb1d(prefix, c):
        conv(prefix + ".conv3x1_1", c, c, 3, 1)
        conv(prefix + ".conv1x3_1", c, c, 1, 3)
        bn(prefix + ".bn1", c)
        conv(prefix + ".conv3x1_2", c, c, 3, 1)
        conv(prefix + ".conv1x3_2", c, c, 1, 3)
        bn(prefix + ".bn2", c)

    def up(prefix, ci, co):
        convT(prefix + ".conv", ci, co, 3, 3)
        bn(prefix + ".bn", co)

    down("encoder.initial_block", in_channels, 16)
    down("encoder.layers.0", 16, 64)
    for i in range(1, 6):
        nb1d("encoder.layers.%d" % i, 64)
    down("encoder.layers.6", 64, 128)
    for i in range(7, 15):
        nb1d("encoder.layers.%d" % i, 128)
    conv("encoder.output_conv", nclasses, 128, 1, 1)
    up("decoder.layers.0", 128, 64)
    nb1d("decoder.layers.1", 64)
    nb1d("decoder.layers.2", 64)
    up("decoder.layers.3", 64, 16)
    nb1d("decoder.layers.4", 16)
    nb1d("decoder.layers.5", 16)
    convT("decoder.output_conv", 16, nclasses, 2, 2)
    return out


def make_erfnet_params(in_channels=3, nclasses=2, seed=0, prefix="net.", bias_scale=0.05):
    """Kaiming-style synthetic parameters keyed like the reference state_dict.

    Same distribution family as ``define_init_weights(model,'kaiming')``
    (Backprojection_Loss/Networks/utils.py:530-543: conv ~ N(0, 2/fan_in), BN weight
    ~ N(1, 0.02)) but with NON-zero conv/BN biases (N(0, bias_scale)) so that bias
    paths are exercised by the parity tests.  fan_in is torch's: size(1)*kh*kw.
    """
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape, kind in erfnet_param_shapes(in_channels, nclasses):
        if kind in ("conv", "convT"):
            fan_in = shape[1] * shape[2] * shape[3]
            w = rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)
        elif kind == "bn_w":
            w = 1.0 + 0.02 * rng.standard_normal(shape)
        else:  # bias, bn_b
            w = bias_scale * rng.standard_normal(shape)
        params[prefix + name] = w.astype(np.float32)
    return params


HEAD_SHAPES = {   # Classification(class_type, size=(32, 64), channels_in=128, resize=256), BP/Networks/LSQ_layer.py:157-192
    "common": [("conv1.weight", (128, 128, 1, 1)), ("conv1.bias", (128,)), ("conv1_bn.weight", (128,)), ("conv1_bn.bias", (128,)),
               ("conv2.weight", (128, 128, 3, 3)), ("conv2.bias", (128,)), ("conv2_bn.weight", (128,)), ("conv2_bn.bias", (128,)),
               ("conv3.weight", (64, 128, 3, 3)), ("conv3.bias", (64,)), ("conv3_bn.weight", (64,)), ("conv3_bn.bias", (64,)),
               ("conv4.weight", (64, 64, 3, 3)), ("conv4.bias", (64,)), ("conv4_bn.weight", (64,)), ("conv4_bn.bias", (64,))],
    "line": [("fully_connected1.weight", (128, 32768)), ("fully_connected1.bias", (128,)),
             ("fully_connected_line1.weight", (4, 128)), ("fully_connected_line1.bias", (4,))],
    "horizon": [("fully_connected_horizon.weight", (256, 2048)), ("fully_connected_horizon.bias", (256,))],
}


def make_head_params(class_type, seed):
    """Synthetic parameters of one Classification head keyed like its state_dict (kaiming-style weights, N(1, 0.02)
    BatchNorm weights, N(0, 0.05) biases so that every bias path is exercised)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape in HEAD_SHAPES["common"] + HEAD_SHAPES[class_type]:
        if name.endswith("_bn.weight"):
            w = 1.0 + 0.02 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            w = 0.05 * rng.standard_normal(shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            w = rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)
        params[name] = w.astype(np.float32)
    return params


def make_encoder_map(B, seed, C=128, H=32, W=64):
    """A post-ReLU-like encoder output [B,C,H,W]: max(0, N(0.3, 1))."""
    rng = np.random.default_rng(seed)
    return np.maximum(rng.standard_normal((B, C, H, W)) + 0.3, 0.0).astype(np.float32)
