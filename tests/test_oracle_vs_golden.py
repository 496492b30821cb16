"""Pin the oracle restatement against outputs of the reference itself
(tests/golden/*.npz, made by oracle/make_golden.py in the build container)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import inputs, lsq_oracle as lo, erfnet_oracle as eo
from oracle.make_golden import LSQ_CASES, NET_CASES, lsq_case_inputs

from conftest import GOLDEN


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def normwise(a, b):
    """max|a-b| / max|b| per coefficient vector."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (np.abs(a - b).max(-1) / np.abs(b).max(-1)).max()


def test_homography_matches_cv2():
    g = load("homography")
    for r in (256, 320):
        M, Mi = lo.get_homography(r)
        # cv2 4.13's getPerspectiveTransform leaves a 1.5e-5 px residual on its own control
        # points (ours: 6e-14), so agreement is ~1e-6 relative, not round-off level.
        np.testing.assert_allclose(M, g["M_%d" % r], rtol=5e-6, atol=1e-9)
        np.testing.assert_allclose(Mi, g["Minv_%d" % r], rtol=5e-6, atol=1e-9)
    M, Mi = lo.get_homography_bev()
    np.testing.assert_allclose(M.astype(np.float32), g["M_bev_f32"], rtol=5e-6, atol=1e-6)


@pytest.mark.parametrize("variant", ["bp", "bev"])
def test_grid_matches_reference(variant):
    if variant == "bp":
        g = load("lsq_bp_l2_d2")
        M = load("homography")["M_256"].astype(np.float32)
        grid = lo.projective_grid(256, 512, M, torch.float32)
    else:
        g = load("lsq_bev_l2_d2")
        M = load("homography")["M_bev_f32"]
        grid = lo.projective_grid(256, 512, M, torch.float32, normalised=True)
    ref = g["grid0"]
    fin = np.isfinite(ref).all(1)
    assert fin.sum() > 0.99 * ref.shape[0]
    np.testing.assert_allclose(grid.numpy()[fin], ref[fin], rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("case", LSQ_CASES, ids=[c[0] for c in LSQ_CASES])
def test_lsq_oracle_matches_reference(case):
    name, variant, L, order, mask_pct, act, reg_ls, chol, maps, B = case
    g = load("lsq_" + name)
    o_np, g_np = lsq_case_inputs(name, L, order, maps, B)
    assert inputs.sha256_of(o_np, g_np) == str(g["input_sha"])
    bev = variant != "Backprojection_Loss"
    gridname = "lsq_bev_l2_d2" if bev else "lsq_bp_l2_d2"
    grid = torch.from_numpy(load(gridname)["grid0"])
    const = 1.0 if bev else 255.0
    zero_rows = lo.mask_rows(256, mask_pct)
    for tag, dt, tol_b, tol_g in (("f64", torch.float64, 1e-9, 1e-7), ("f32", torch.float32, 3e-4, None)):
        o = torch.from_numpy(o_np).to(dt).requires_grad_(True)
        masked = lo.activate_and_mask(o, act, zero_rows)
        beta, Zinv = lo.wls_forward(masked, grid, order, const, reg_ls, use_cholesky=chol)
        ref = g["beta_" + tag]
        assert normwise(beta.detach().numpy(), ref) < tol_b, (tag, normwise(beta.detach().numpy(), ref))
        if tol_g is None:
            continue
        gb = torch.from_numpy(g_np)
        (beta * gb).sum().backward()
        idx, val, stat = g["grad_f64/idx"], g["grad_f64/val"], g["grad_f64/stat"]
        got = o.grad.numpy().reshape(-1)
        assert np.abs(got[idx] - val).max() <= tol_g * stat[2]
        # closed-form backward (what the CUDA kernel implements) == autograd of the reference
        cf = lo.wls_backward_closed_form(o_np, grid.numpy(), order, beta.detach().numpy(), Zinv.detach().numpy(),
                                         g_np, act, zero_rows, const).reshape(-1)
        assert np.abs(cf[idx] - val).max() <= 1e-7 * stat[2]
        assert abs(np.sqrt((cf * cf).sum()) - stat[3]) <= 1e-7 * stat[3]


@pytest.mark.parametrize("case", NET_CASES, ids=[c[0] for c in NET_CASES])
def test_full_path_oracle_matches_reference(case):
    name, L, order, mask_pct, B = case
    g = load(name)
    meta = json.loads(str(g["meta"]))
    P_np = inputs.make_erfnet_params(3, L, seed=meta["param_seed"])
    assert inputs.sha256_of(*[P_np[k] for k in sorted(P_np)]) == str(g["param_sha"])
    x_np = inputs.make_images(B, 256, 512, seed=meta["image_seed"])
    xgt_np, valid_np = inputs.make_loss_targets(B, 4, seed=meta["target_seed"])
    assert inputs.sha256_of(x_np, xgt_np, valid_np) == str(g["input_sha"])
    grid = torch.from_numpy(load("lsq_bp_l2_d2")["grid0"])
    hg = load("homography")            # the reference's own cv2 matrices (see test_homography_matches_cv2)
    crit = lo.BackprojectionLoss(order, 256, M=hg["M_256"], M_inv=hg["Minv_256"])
    for tag, dt, tol in (("f64", torch.float64, 1e-8), ("f32", torch.float32, 2e-3)):
        P = {k[4:]: torch.from_numpy(v).to(dt).requires_grad_(True) for k, v in P_np.items()}
        taps, stats = {}, {}
        loss, beta, dec, masked = eo.full_step(
            torch.from_numpy(x_np).to(dt), P, grid, order, L, lo.mask_rows(256, mask_pct),
            torch.from_numpy(xgt_np), torch.from_numpy(valid_np), taps=taps, stats_out=stats, loss_obj=crit)
        loss.backward()
        assert abs(float(loss.detach()) - float(g["loss_" + tag])) <= tol * abs(float(g["loss_" + tag]))
        assert normwise(beta.detach().numpy(), g["beta_" + tag]) < tol
        for n, t in taps.items():
            key = "act_%s/%s" % (tag, n)
            got = t.detach().double().numpy().reshape(-1)[g[key + "/idx"]]
            assert np.abs(got - g[key + "/val"]).max() <= tol * g[key + "/stat"][2], n
        if tag == "f64":
            # conv biases that feed a BatchNorm have an analytically zero gradient (round-off
            # noise only), so the per-tensor relative gate gets a global absolute floor.
            gscale = max(g[k][2] for k in g.files if k.startswith("grad_f64/") and k.endswith("/stat"))
            for n, p in P.items():
                key = "grad_f64/net." + n
                if key + "/idx" not in g:
                    assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
                    continue
                got = p.grad.numpy().reshape(-1)[g[key + "/idx"]]
                assert np.abs(got - g[key + "/val"]).max() <= 1e-7 * g[key + "/stat"][2] + 1e-12 * gscale, n
            # BN running stats after one step: momentum 0.1, unbiased variance
            for n, (mean, var_unb) in stats.items():
                np.testing.assert_allclose(0.1 * mean.numpy(), g["buf_f64/net.%s.running_mean" % n], rtol=1e-9, atol=1e-12)
                np.testing.assert_allclose(0.9 + 0.1 * var_unb.numpy(), g["buf_f64/net.%s.running_var" % n], rtol=1e-9)


@pytest.mark.parametrize("order,wf", [(2, "none"), (2, "linear"), (2, "quadratic"), (1, "none")])
def test_area_loss_matches_reference_golden(order, wf):
    """A12: the oracle's bool-mask restatement AND the product's Area_Loss (pure torch ops, any device) against the golden
    outputs of the reference's own Area_Loss (BP/Loss_crit.py:87-143 == BEV/Loss_crit.py:78-134, run with the torch-1.1
    byte-mask dispatch shimmed: oracle/make_golden.py run_area_loss) -- value and gradient, fp64 to 1e-12, fp32 to 1e-6."""
    import numpy as np
    from lanedetection_end2end_b200.Loss_crit import Area_Loss
    g = np.load(os.path.join(GOLDEN, "area_loss.npz"))
    for variant in ("Ba", "Bi"):
        for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 2e-6)):
            key = "%s/o%d_%s_%s" % (variant, order, wf, tag)
            want, wgrad = float(g[key + "/loss"]), g[key + "/grad"]
            for impl in ("oracle", "product"):
                p = torch.from_numpy(g["params"]).to(dt).requires_grad_(True)
                gt = torch.from_numpy(g["gt"]).to(dt)
                loss = lo.area_loss(p, gt, order, wf) if impl == "oracle" else Area_Loss(order, wf)(p, gt)
                loss.backward()
                assert abs(float(loss.detach()) - want) <= tol * abs(want), (impl, key)
                assert np.abs(p.grad.double().numpy() - wgrad).max() <= tol * np.abs(wgrad).max(), (impl, key)
        z = Area_Loss(2, "none")(torch.from_numpy(g["params"]).float(), torch.zeros(6, 3))
        assert float(z) == float(g[variant + "/all_absent"]) == 0.0


@pytest.mark.parametrize("kind", ["line", "horizon"])
def test_heads_restatement_matches_reference_golden(kind):
    """oracle/heads_oracle.py (fp64) against the reference's own Classification class (clas_heads.npz): forward and the
    sampled input / parameter gradients."""
    import json
    from oracle import heads_oracle
    g = np.load(os.path.join(GOLDEN, "clas_heads.npz"))
    meta = json.loads(str(g["meta"]))
    P = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in inputs.make_head_params(kind, meta["param_seeds"][kind]).items()}
    x = torch.from_numpy(inputs.make_encoder_map(meta["B"], seed=meta["map_seed"])).double().requires_grad_(True)
    y = heads_oracle.head_forward(P, x, kind)
    ref = g["%s/out_f64" % kind]
    assert np.abs(y.detach().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
    (y * torch.from_numpy(g["%s/g" % kind])).sum().backward()
    k = "%s/dx_f64" % kind
    assert np.abs(x.grad.numpy().reshape(-1)[g[k + "/idx"]] - g[k + "/val"]).max() <= 1e-8 * g[k + "/stat"][2]
    for n, p in P.items():
        k = "%s/grad_f64/%s" % (kind, n)
        scale = max(g[k + "/stat"][2], 1e-12)
        if n.startswith("conv") and n.endswith(".bias") and "_bn" not in n:
            continue        # analytically zero (conv feeding a training-mode BatchNorm): pure round-off on both sides
        assert np.abs(p.grad.numpy().reshape(-1)[g[k + "/idx"]] - g[k + "/val"]).max() <= 1e-7 * scale, n
