"""GPU parity of the Classification heads (`--clas 1`; BP/Networks/LSQ_layer.py:157-207) on the library's kernels against
the committed golden outputs of the reference's own class (tests/golden/clas_heads.npz, oracle/make_golden.py
run_clas_heads): forward, input gradient, every parameter gradient and the BatchNorm running statistics, in both
fp32-accurate convolution modes.  Gate: |ours - fp64| <= 4 |reference fp32 - fp64| + tol, norm-wise.
Plus the fully connected kernels alone against torch-CPU float64 at awkward sizes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import inputs
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(params=["tf32x3", "fp32"])
def conv_mode(request):
    from lanedetection_end2end_b200 import ops_net
    ops_net.set_conv_mode(request.param)
    yield request.param


def _head(kind):
    from lanedetection_end2end_b200.Networks.LSQ_layer import Classification
    g = np.load(os.path.join(GOLDEN, "clas_heads.npz"))
    meta = json.loads(str(g["meta"]))
    m = Classification(kind, size=(32, 64), channels_in=128, resize=256)
    sd = m.state_dict()
    for k, v in inputs.make_head_params(kind, meta["param_seeds"][kind]).items():
        assert tuple(sd[k].shape) == v.shape, k
        sd[k] = torch.from_numpy(v)
    m.load_state_dict(sd)
    return m.cuda().train(), g, meta


@pytest.mark.parametrize("kind", ["line", "horizon"])
def test_head_matches_reference_golden(kind, conv_mode):
    m, g, meta = _head(kind)
    x = torch.from_numpy(inputs.make_encoder_map(meta["B"], seed=meta["map_seed"])).cuda().requires_grad_(True)
    y = m(x)
    assert y.shape == g["%s/out_f64" % kind].shape and y.dtype == torch.float32
    (y * torch.from_numpy(g["%s/g" % kind]).float().cuda()).sum().backward()
    torch.cuda.synchronize()

    def gate(name, ours, k64, k32, tol=TOL):
        ref64, ref32 = g[k64], g[k32]
        scale = max(float(np.abs(ref64).max()), 1e-30)
        e_ours = float(np.abs(ours - ref64).max() / scale)
        e_ref = float(np.abs(ref32 - ref64).max() / scale)
        assert e_ours <= 4 * e_ref + tol, (kind, conv_mode, name, e_ours, e_ref)
        return e_ours, e_ref

    gate("out", y.detach().double().cpu().numpy(), "%s/out_f64" % kind, "%s/out_f32" % kind)
    k = "%s/dx_f64" % kind
    got = x.grad.double().cpu().numpy().reshape(-1)[g[k + "/idx"]]
    scale = g[k + "/stat"][2]
    e_ours = float(np.abs(got - g[k + "/val"]).max() / scale)
    e_ref = float(np.abs(g["%s/dx_f32/val" % kind] - g[k + "/val"]).max() / scale)
    assert e_ours <= 4 * e_ref + 10 * TOL, (kind, "dx", e_ours, e_ref)
    gscale = max(g[q][2] for q in g.files if q.startswith("%s/grad_f64/" % kind) and q.endswith("/stat"))
    for n, p in m.named_parameters():
        k64, k32 = "%s/grad_f64/%s" % (kind, n), "%s/grad_f32/%s" % (kind, n)
        got = p.grad.double().cpu().numpy().reshape(-1)[g[k64 + "/idx"]]
        scale = max(g[k64 + "/stat"][2], 1e-6 * gscale)
        e_ours = float(np.abs(got - g[k64 + "/val"]).max() / scale)
        e_ref = float(np.abs(g[k32 + "/val"] - g[k64 + "/val"]).max() / scale)
        assert e_ours <= 4 * e_ref + 10 * TOL, (kind, conv_mode, n, e_ours, e_ref)
    for n, b in m.named_buffers():
        if n.endswith("running_mean") or n.endswith("running_var"):
            ref = g["%s/buf_f64/%s" % (kind, n)]
            assert np.abs(b.cpu().numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), (kind, n)
        if n.endswith("num_batches_tracked"):
            assert int(b) == 1


@pytest.mark.parametrize("B,K,O,relu", [(2, 32768, 128, True), (5, 128, 4, False), (33, 2048, 256, False), (70, 700, 12, True)])
def test_linear_kernels_vs_torch_fp64(B, K, O, relu):
    from lanedetection_end2end_b200.ops_heads import LinearFunction
    gen = torch.Generator().manual_seed(B * 1000 + O)
    x = torch.randn(B, K, generator=gen)
    W = torch.randn(O, K, generator=gen) / K ** 0.5
    b = torch.randn(O, generator=gen) * 0.1
    gy = torch.randn(B, O, generator=gen)
    xd, Wd, bd = (t.double().requires_grad_(True) for t in (x, W, b))
    yd = xd @ Wd.t() + bd
    yd = yd.relu() if relu else yd
    (yd * gy.double()).sum().backward()
    xg, Wg, bg = (t.cuda().requires_grad_(True) for t in (x, W, b))
    yg = LinearFunction.apply(xg, Wg, bg, relu)
    (yg * gy.cuda()).sum().backward()
    torch.cuda.synchronize()
    for name, a, r in (("y", yg, yd), ("dx", xg.grad, xd.grad), ("dW", Wg.grad, Wd.grad), ("db", bg.grad, bd.grad)):
        err = float((a.detach().double().cpu() - r.detach()).abs().max() / r.detach().abs().max().clamp_min(1e-30))
        assert err < 2e-5, (name, B, K, O, err)


def test_net_with_clas_returns_heads_and_backpropagates():
    """`--clas 1` through Net.forward (BP/Networks/LSQ_layer.py:250-257,296-298): line [B,4], horizon [B,resize], both
    differentiable down to the encoder; the LSQ outputs are unchanged by the heads."""
    from lanedetection_end2end_b200.Networks.utils import define_args
    from lanedetection_end2end_b200.Networks.LSQ_layer import Net
    B = 2
    args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--nclasses", "4", "--order", "3", "--batch_size", str(B),
                                     "--mask_percentage", "0.2", "--loss_policy", "backproject", "--clas", "1"])
    torch.manual_seed(0)
    model = Net(args).cuda().train()
    x = torch.from_numpy(inputs.make_images(B, 256, 512, seed=3)).cuda()
    out = model(x, torch.zeros(B, 4), True)
    line, horizon = out[6], out[7]
    assert line.shape == (B, 4) and horizon.shape == (B, 256)
    (line.sum() + horizon.sum() + sum(b.sum() for b in out[:4]).float() * 0).backward()
    torch.cuda.synchronize()
    assert model.line_classification.fully_connected1.weight.grad is not None
    gin = model.net.encoder.initial_block.conv.weight.grad
    assert gin is not None and torch.isfinite(gin).all() and float(gin.abs().max()) > 0
