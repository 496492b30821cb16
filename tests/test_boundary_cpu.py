"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares (no compute calls), the reference-mirroring modules keep names / state_dict keys / flag defaults /
error conventions, and device ops refuse CPU tensors loudly (no fallback)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def test_c_abi_library_exports_every_declared_symbol():
    from lanedetection_end2end_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "run __graft_entry__.build() first"
    hdr = open(os.path.join(ROOT, "include", "lanefit_b200.h")).read()
    declared = set(re.findall(r"\b(lf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    h = _capi.lib()
    for name in declared:
        assert hasattr(h, name), name
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    # host-only entry points are callable without a GPU
    assert h.lf_version() >= 100
    assert h.lf_error_string(-2) == b"unsupported configuration"
    assert h.lf_lsq_workspace_bytes(2, 2, 256, 512, 2) > 0
    assert h.lf_bn_blocks(65536, 128) >= 148
    assert h.lf_lsq_fwd(None, 0, None, None, None, 1, 1, 8, 8, 2, 0, 1, 0.0, 0, None, None, None, None, None, 0, None) == -1


def test_ctypes_struct_layout_matches_header():
    """Field order / count of the ctypes mirrors vs the C structs in the header."""
    from lanedetection_end2end_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "lanefit_b200.h")).read()
    for cname, cls in (("LfConvArgs", _capi.LfConvArgs), ("LfWgradArgs", _capi.LfWgradArgs),
                       ("LfConvTcArgs", _capi.LfConvTcArgs), ("LfTcgView", _capi.LfTcgView),
                       ("LfConvTcgArgs", _capi.LfConvTcgArgs), ("LfReduceJob", _capi.LfReduceJob),
                       ("LfWgradTcgArgs", _capi.LfWgradTcgArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(float|int|double|long long|LfTcgView)\s*\*?", "", decl)
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part.replace("*", "")).strip())
        mine = [f[0] for f in cls._fields_]
        assert len(names) == len(mine), (cname, names, mine)
        for a, b in zip(names, mine):
            assert a == b or (a == "in" and b == "inp"), (cname, a, b)


def test_state_dict_keys_match_reference():
    from lanedetection_end2end_b200.Networks.utils import define_args
    from lanedetection_end2end_b200.Networks.LSQ_layer import Net
    for name, L in (("net_l2_d2", 2), ("net_l4_d3", 4)):
        ref = json.loads(str(np.load(os.path.join(GOLDEN, name + ".npz"))["state_dict_keys"]))
        args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--no_cuda", "--nclasses", str(L),
                                         "--batch_size", "2"])
        m = Net(args)
        mine = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
        assert mine == ref
        # plain attributes, not buffers (BP/Networks/LSQ_layer.py:77-83,231,238)
        assert not any(k.startswith(("grid", "ls_layer", "idx_row")) for k in m.state_dict())
        assert tuple(m.grid.shape) == (2, 256 * 512, 2)


def test_grid_is_bit_identical_to_reference():
    from lanedetection_end2end_b200.Networks.LSQ_layer import ProjectiveGridGenerator
    from lanedetection_end2end_b200.Networks.utils import get_homography
    M, Minv = get_homography(256)
    g = np.load(os.path.join(GOLDEN, "homography.npz"))
    np.testing.assert_array_equal(M, g["M_256"])          # same cv2 -> same matrix
    np.testing.assert_array_equal(Minv, g["Minv_256"])
    grid = ProjectiveGridGenerator(torch.Size([3, 2, 256, 512]), torch.from_numpy(M).float().unsqueeze(0).expand(3, 3, 3), True)
    ref = np.load(os.path.join(GOLDEN, "lsq_bp_l2_d2.npz"))["grid0"]
    np.testing.assert_array_equal(grid[0].numpy(), ref)
    np.testing.assert_array_equal(grid[2].numpy(), ref)


def test_flag_defaults_and_required_flags():
    from lanedetection_end2end_b200.Networks.utils import define_args
    p = define_args()
    with pytest.raises(SystemExit):
        p.parse_args([])                                   # --image_dir / --gt_dir are required
    a = p.parse_args(["--image_dir", "x", "--gt_dir", "y"])
    assert (a.nclasses, a.order, a.resize, a.batch_size, a.mask_percentage) == (2, 2, 256, 8, 0.3)
    assert a.activation_layer == "square" and a.reg_ls == 0 and a.use_cholesky is False and a.end_to_end is True
    assert a.weight_init == "kaiming" and a.loss_policy == "area" and a.mod == "erfnet" and a.clas is False
    assert p.parse_args(["--image_dir", "x", "--gt_dir", "y", "--use_cholesky", "1", "--end_to_end", "False"]).use_cholesky
    with pytest.raises(SystemExit):
        p.parse_args(["--image_dir", "x", "--gt_dir", "y", "--nclasses", "3"])


def test_error_conventions_and_no_cpu_fallback():
    from lanedetection_end2end_b200 import Networks
    from lanedetection_end2end_b200.Networks import LSQ_layer, ERFNet
    from lanedetection_end2end_b200._capi import LanefitError
    with pytest.raises(KeyError):
        Networks.define_model("segnet")
    with pytest.raises(NotImplementedError):
        LSQ_layer.activation_layer("tanh")
    blk = ERFNet.non_bottleneck_1d(16, 0.0, 1)
    with pytest.raises(LanefitError):
        blk(torch.randn(1, 16, 8, 8))                      # CPU tensor: refused, never silently computed
    ls = LSQ_layer.Weighted_least_squares(torch.Size([1, 2, 8, 16]), 2, 2, True)
    with pytest.raises(LanefitError):
        ls(torch.rand(1, 2, 8, 16), torch.rand(1, 128, 2))


def test_loss_modules_match_oracle_on_cpu():
    """The losses are O(B*56) float64 torch code and run anywhere; check them against the oracle."""
    from types import SimpleNamespace
    from lanedetection_end2end_b200.Loss_crit import backprojection_loss, Area_Loss
    from oracle import lsq_oracle as lo
    g = np.load(os.path.join(GOLDEN, "homography.npz"))
    opt = SimpleNamespace(resize=256, no_mapping=False, order=2, batch_size=4, no_cuda=True)
    crit = backprojection_loss(opt)
    ref = lo.BackprojectionLoss(2, 256, M=g["M_256"], M_inv=g["Minv_256"])
    beta = torch.tensor([[1e-3, -0.2, 260.0], [2e-3, 0.1, 250.0], [0.0, 0.0, 270.0], [-1e-3, 0.3, 240.0]],
                        dtype=torch.float64).unsqueeze(-1).requires_grad_(True)
    xgt = torch.rand(4, 56, dtype=torch.float64) * 500
    valid = torch.ones(4, 56, dtype=torch.float64)
    valid[:, :8] = 0
    l1, x1 = crit(beta, xgt, valid)
    l2, x2 = ref(beta.detach(), xgt, valid)
    assert abs(float(l1.detach()) - float(l2)) <= 1e-12 * abs(float(l2))
    torch.testing.assert_close(x1.detach(), x2, rtol=1e-12, atol=1e-9)
    l1.backward()
    assert torch.isfinite(beta.grad).all()
    gt = torch.tensor([[1e-3, -0.2, 0.5], [0.0, 0.0, 0.0], [2e-3, 0.1, 0.4], [1e-3, 0.2, 0.6]], dtype=torch.float64)
    for wf in ("none", "linear", "quadratic"):
        a = Area_Loss(2, wf)(beta.detach() * 1e-3, gt)
        b = lo.area_loss(beta.detach() * 1e-3, gt, 2, wf)
        assert abs(float(a) - float(b)) <= 1e-12 * max(1.0, abs(float(b)))


def test_reference_main_imports_resolve_with_package_dir_on_path():
    """`from Networks.LSQ_layer import Net`, `from Loss_crit import define_loss_crit` ... -- the imports at
    the top of the reference's main.py (BP/main.py:24-28) resolve against this package directory."""
    code = ("import sys; sys.path.insert(0, %r);"
            "from Loss_crit import define_loss_crit, backprojection_loss;"
            "from Networks.LSQ_layer import Net;"
            "from Networks.utils import define_args, save_weightmap, first_run, mkdir_if_missing, Logger, "
            "define_init_weights, define_scheduler, define_optim, AverageMeter;"
            "from Networks.gels import GELS; import Networks; print(sorted(Networks.model_dict))")
    pkg = os.path.join(ROOT, "lanedetection_end2end_b200")
    out = subprocess.run([sys.executable, "-c", code % pkg], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert "erfnet" in out.stdout


def test_fused_dropout_masks_statistics_and_layout():
    """One uniform draw feeds the Dropout2d channel masks of every block (Networks/ERFNet.py: fused_dropout_masks):
    right shapes, values in {0, 1/(1-p)}, keep rate ~ 1-p per block, reproducible under the global seed, and each
    block consumes its mask exactly once."""
    import torch
    from lanedetection_end2end_b200.Networks import ERFNet as E
    blocks = [E.non_bottleneck_1d(64, 0.03, 1) for _ in range(3)] + [E.non_bottleneck_1d(128, 0.3, 2) for _ in range(4)]
    B = 64
    torch.manual_seed(7)
    masks = E.fused_dropout_masks(blocks, B, torch.device("cpu"))
    torch.manual_seed(7)
    again = E.fused_dropout_masks(blocks, B, torch.device("cpu"))
    for b, m, m2 in zip(blocks, masks, again):
        c, p = b.conv3x1_1.out_channels, b.dropout.p
        assert m.shape == (B, c) and m.is_contiguous() and torch.equal(m, m2)
        vals = set(torch.unique(m).tolist())
        assert vals <= {0.0, float(torch.tensor(1.0 / (1.0 - p), dtype=torch.float32))}
        keep = float((m > 0).float().mean())
        assert abs(keep - (1 - p)) < 4 * (p * (1 - p) / (B * c)) ** 0.5 + 1e-3, (p, keep)
        x = torch.zeros(B, 4, 4, c)                       # NHWC like the block sees it
        b.train()
        got = b._drop_mask(x)
        assert got is m2                                   # the pending slice is handed out ...
        fresh = b._drop_mask(x)
        assert fresh is not m2 and fresh.shape == (B, c)   # ... exactly once; afterwards the block draws its own


def test_fused_backprojection_loss_math_on_host():
    """csrc/loss.cu: the per-point code of the fused multi-lane back-projection loss kernel is __host__ __device__;
    lf_backproj_loss_host runs it on the CPU.  Pin loss, x_cal and d loss / d beta against the per-lane torch module
    (which the oracle tests pin against the reference) -- including a lane without any valid sample."""
    import ctypes
    import numpy as np
    import torch
    from lanedetection_end2end_b200 import _capi
    from lanedetection_end2end_b200.Loss_crit import backprojection_loss
    from lanedetection_end2end_b200.Networks.utils import define_args
    for order, L, B in ((2, 2, 5), (3, 4, 3), (1, 3, 2)):
        args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--no_cuda", "--order", str(order),
                                         "--nclasses", "4" if L > 2 else "2"])
        crit = backprojection_loss(args)
        g = torch.Generator().manual_seed(order * 10 + L)
        n = order + 1
        beta = (torch.randn(B, L, n, generator=g, dtype=torch.float64) * torch.tensor([1e-3, 0.1, 100.0, 1.0][-n:])).requires_grad_(True)
        x_gt = torch.rand(B, L, 56, generator=g, dtype=torch.float64) * 500
        valid = (torch.rand(B, L, 56, generator=g) > 0.3).double()
        valid[:, L - 1] = 0                                   # last lane: nothing valid -> contributes 0, finite gradient 0
        total = 0
        xcals = []
        for l in range(L):
            ll, xc = crit(beta[:, l].unsqueeze(-1), x_gt[:, l], valid[:, l])
            total = total + ll
            xcals.append(xc)
        ref = total / L
        ref.backward()
        Y, yp, Mi = crit._fused_host_constants()
        h = _capi.lib()
        lane = np.zeros(L)
        loss = np.zeros(1)
        dbeta = np.zeros((B, L, n))
        xcal = np.zeros((B, L, 56))
        bt = np.ascontiguousarray(beta.detach().numpy())
        P = lambda a: ctypes.c_void_p(a.ctypes.data)
        rc = h.lf_backproj_loss_host(P(Y), P(yp), P(Mi), P(bt), P(np.ascontiguousarray(x_gt.numpy())),
                                     P(np.ascontiguousarray(valid.numpy())), B, L, n, P(lane), P(loss), P(dbeta), P(xcal))
        assert rc == 0
        assert abs(loss[0] - float(ref.detach())) <= 1e-12 * abs(float(ref.detach()))
        np.testing.assert_allclose(xcal, torch.stack(xcals, 1).detach().numpy(), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(dbeta, beta.grad.numpy(), rtol=1e-9, atol=1e-12 * np.abs(beta.grad.numpy()).max())
        assert lane[L - 1] == 0.0 and not np.any(dbeta[:, L - 1])


def test_backprojection_loss_all_lanes_at_once_equals_the_lane_loop():
    """backprojection_loss.forward_lanes == the reference's per-lane loop (BP/main.py:297-305): value, x_cal, gradients."""
    import torch
    from lanedetection_end2end_b200.Loss_crit import backprojection_loss
    from lanedetection_end2end_b200.Networks.utils import define_args
    for order, L, B in ((2, 2, 6), (3, 4, 3)):
        args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--no_cuda", "--order", str(order),
                                         "--nclasses", str(L)])
        crit = backprojection_loss(args)
        g = torch.Generator().manual_seed(order)
        n = order + 1
        scale = torch.tensor([1e-3, 0.1, 100.0, 1.0][-n:], dtype=torch.float64).view(1, n, 1)
        betas = [(torch.randn(B, n, 1, generator=g, dtype=torch.float64) * scale).requires_grad_(True) for _ in range(L)]
        x_gt = torch.rand(B, 4, 56, generator=g, dtype=torch.float64) * 500
        valid = (torch.rand(B, 4, 56, generator=g) > 0.3).double()
        valid[:, 1] = 0
        ref = sum(crit(betas[l], x_gt[:, l], valid[:, l])[0] for l in range(L)) / L
        gref = torch.autograd.grad(ref, betas)
        loss, xcal = crit.forward_lanes(betas, x_gt, valid)
        gours = torch.autograd.grad(loss, betas)
        assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-13 * abs(float(ref.detach()))
        for l in range(L):
            torch.testing.assert_close(xcal[:, l], crit(betas[l], x_gt[:, l], valid[:, l])[1], rtol=1e-13, atol=1e-10)
            torch.testing.assert_close(gours[l], gref[l], rtol=1e-10, atol=1e-16)


def test_classification_heads_state_dict_and_no_cpu_fallback():
    """`--clas 1` heads (BP/Networks/LSQ_layer.py:157-207): parameter names / shapes of the reference (checkpoints round-trip),
    and no CPU path -- a CPU tensor raises instead of silently running torch modules."""
    import pytest
    import torch
    from oracle import inputs
    from lanedetection_end2end_b200 import _capi
    from lanedetection_end2end_b200.Networks.LSQ_layer import Classification
    for kind in ("line", "horizon"):
        m = Classification(kind, size=(32, 64), channels_in=128, resize=256)
        sd = m.state_dict()
        want = dict(inputs.HEAD_SHAPES["common"] + inputs.HEAD_SHAPES[kind])
        have = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
        assert have == want
        assert m.conv1_bn.eps == 1e-5
        with pytest.raises(_capi.LanefitError):
            m(torch.zeros(1, 128, 32, 64))


def test_tusimple_writer_and_lane_gating_host_logic():
    """inference.lanes_from_predictions / write_tusimple_predictions (BP/test.py:66-99) are host logic: lanes switched off by
    the line-type head, samples above the horizon and outside the frame become -2; one JSON line per image."""
    import io
    import json
    import torch
    from lanedetection_end2end_b200.inference import lanes_from_predictions, write_tusimple_predictions
    x = torch.full((2, 4, 56), 100.0, dtype=torch.float64)
    x[0, 0, 10] = 1500.0            # outside the 1280-wide frame
    x[1, 2, 5] = -3.0
    lanes = lanes_from_predictions(x, torch.tensor([[1., 1, 0, 1], [1., 1, 1, 1]]), torch.tensor([200, 160]))
    assert lanes[0][0][10] == -2 and lanes[1][2][5] == -2
    assert all(v == -2 for v in lanes[0][1])                 # head order [1, 2, 0, 3]: output lane 1 <- line_pred[:, 2] = 0 -> gated
    assert lanes[0][0][:4] == [-2] * 4 and lanes[0][0][4] == 100    # horizon 200 -> the first (200 - 160) / 10 samples
    assert lanes[1][0][0] == 100
    buf = io.StringIO()
    gt = [{"raw_file": "a.jpg", "h_samples": list(range(160, 720, 10))}, {"raw_file": "b.jpg", "h_samples": list(range(160, 720, 10))}]
    write_tusimple_predictions(buf, gt, lanes, 0)
    rows = [json.loads(l) for l in buf.getvalue().splitlines()]
    assert len(rows) == 2 and rows[1]["raw_file"] == "b.jpg" and rows[0]["run_time"] == 20 and rows[0]["lanes"] == lanes[0]


def test_eval_batchnorm_folding_algebra():
    """ops_eval folds an eval-mode BatchNorm into the convolution that feeds it (w' = w * scale[co], b' = b * scale + shift;
    ConvTranspose2d: scale along dim 1).  The algebra, checked with torch's own convolutions on the CPU (the kernels are
    checked on the GPU by test_eval_fused_matches_unfused...)."""
    import torch
    import torch.nn.functional as F
    from lanedetection_end2end_b200 import ops_eval
    g = torch.Generator().manual_seed(0)
    bn = torch.nn.BatchNorm2d(8, eps=1e-3).double().eval()
    with torch.no_grad():
        bn.weight.copy_(torch.randn(8, generator=g, dtype=torch.float64))
        bn.bias.copy_(torch.randn(8, generator=g, dtype=torch.float64))
        bn.running_mean.copy_(torch.randn(8, generator=g, dtype=torch.float64))
        bn.running_var.copy_(torch.rand(8, generator=g, dtype=torch.float64) + 0.5)
    scale, shift = ops_eval.bn_affine(bn)
    x = torch.randn(2, 5, 9, 11, generator=g, dtype=torch.float64)
    w = torch.randn(8, 5, 3, 1, generator=g, dtype=torch.float64)
    b = torch.randn(8, generator=g, dtype=torch.float64)
    want = bn(F.conv2d(x, w, b, padding=(1, 0)))
    got = F.conv2d(x, w * scale.view(-1, 1, 1, 1), b * scale + shift, padding=(1, 0))
    assert float((got - want).abs().max()) < 1e-12
    wt = torch.randn(5, 8, 3, 3, generator=g, dtype=torch.float64)          # ConvTranspose2d: [Cin, Cout, kh, kw]
    want = bn(F.conv_transpose2d(x, wt, b, stride=2, padding=1, output_padding=1))
    got = F.conv_transpose2d(x, wt * scale.view(1, -1, 1, 1), b * scale + shift, stride=2, padding=1, output_padding=1)
    assert float((got - want).abs().max()) < 1e-12
