"""CPU restatement of the weighted least-squares layer, its inputs and the
losses that consume beta (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Each function cites the reference lines it restates (paths relative to
/root/reference).  All functions are dtype-generic torch-CPU code: run with
float32 they mirror the reference's arithmetic op-for-op; run with float64
they are the arbiter of SURVEY.md section 7.2 #1.  Generalisations beyond the
reference API (any number of lanes, order up to 4) follow the same formulas.
"""
import math

import numpy as np
import torch


# ---------------------------------------------------------------------------
# Homography (Backprojection_Loss/Networks/utils.py:104-121)
# ---------------------------------------------------------------------------

def perspective_transform(src, dst):
    """Restatement of cv2.getPerspectiveTransform (OpenCV imgproc, not vendored in
    the reference): solve the 8x8 DLT system for the homography mapping the four
    ``src`` points onto ``dst`` with h33 = 1, in float64."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    A = np.zeros((8, 8), dtype=np.float64)
    b = np.zeros(8, dtype=np.float64)
    for i in range(4):
        x, y = src[i]
        X, Y = dst[i]
        A[i] = [x, y, 1, 0, 0, 0, -x * X, -y * X]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * Y, -y * Y]
        b[i], b[i + 4] = X, Y
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


def get_homography(resize=256, no_mapping=False):
    """Backprojection_Loss/Networks/utils.py:104-121."""
    if no_mapping:
        return np.identity(3), np.identity(3)
    y_start = 0.20 * resize
    y_stop = resize - 1
    src = np.float32([[0.45 * (2 * resize), y_start], [0.55 * (2 * resize), y_start],
                      [0.02 * (2 * resize), y_stop], [0.97 * (2 * resize), y_stop]])
    dst = np.float32([[0.45 * (2 * resize), y_start], [0.55 * (2 * resize), y_start],
                      [0.45 * (2 * resize), y_stop], [0.55 * (2 * resize), y_stop]])
    return perspective_transform(src, dst), perspective_transform(dst, src)


def get_homography_bev():
    """Birds_Eye_View_Loss/Networks/LSQ_layer.py:17-32 (normalised coordinates)."""
    y_start, y_stop = 0.3, 1
    src = np.float32([[0.45, y_start], [0.55, y_start], [0.1, y_stop], [0.9, y_stop]])
    dst = np.float32([[0.45, y_start], [0.55, y_start], [0.45, y_stop], [0.55, y_stop]])
    return perspective_transform(src, dst), perspective_transform(dst, src)


# ---------------------------------------------------------------------------
# Grid (Backprojection_Loss/Networks/LSQ_layer.py:50-68; BEV :66-87)
# ---------------------------------------------------------------------------

def projective_grid(H, W, theta, dtype=torch.float32, normalised=False):
    """Single-image grid [H*W, 2] = perspective-divided (x, y, 1) @ theta^T.
    ``normalised`` selects the BEV variant's [0,1) base coordinates."""
    theta = torch.as_tensor(theta).to(dtype)
    if normalised:
        lw = torch.linspace(0, 1 - 1 / W, W, dtype=dtype)
        lh = torch.linspace(0, 1 - 1 / H, H, dtype=dtype)
    else:
        lw = torch.linspace(0, W - 1, W, dtype=dtype)
        lh = torch.linspace(0, H - 1, H, dtype=dtype)
    base = torch.empty(H, W, 3, dtype=dtype)
    base[:, :, 0] = torch.outer(torch.ones(H, dtype=dtype), lw)
    base[:, :, 1] = torch.outer(lh, torch.ones(W, dtype=dtype))
    base[:, :, 2] = 1
    g = torch.bmm(base.view(1, H * W, 3), theta.view(1, 3, 3).transpose(1, 2))[0]
    return g[:, 0:2] / g[:, 2:]


# ---------------------------------------------------------------------------
# Activation + row mask (LSQ_layer.py:27-47, :237-238, :295, :301)
# ---------------------------------------------------------------------------

ACTIVATIONS = ("none", "square", "abs", "relu", "sigmoid", "softplus")


def activation(o, kind="square"):
    if kind == "square":
        return o ** 2
    if kind == "abs":
        return torch.abs(o)
    if kind == "relu":
        return torch.relu(o)
    if kind == "sigmoid":
        return torch.sigmoid(o)
    if kind == "softplus":
        return torch.nn.functional.softplus(o)
    if kind == "none":
        return o
    raise NotImplementedError("Activation type: {} is not implemented".format(kind))


def mask_rows(resize, mask_percentage):
    return int(math.ceil(resize * mask_percentage))


def activate_and_mask(o, kind, zero_rows):
    """masked = activation(o).index_fill(2, rows < zero_rows, 0)  (LSQ_layer.py:295,301)."""
    a = activation(o, kind)
    if zero_rows > 0:
        a = a.clone()
        a[:, :, :zero_rows, :] = 0
    return a


# ---------------------------------------------------------------------------
# Weighted least squares (LSQ_layer.py:85-154), generalised over lanes/order
# ---------------------------------------------------------------------------

def design_matrix(y, order):
    """Y = [y^d ... y 1]  (LSQ_layer.py:97-107; highest power first)."""
    cols = [y ** k for k in range(order, 0, -1)] + [torch.ones_like(y)]
    return torch.stack(cols, dim=-1)


def wls_forward(Wmap, grid, order, const=255.0, reg_ls=0.0, skip_rows=0, use_cholesky=False):
    """beta[B,L,order+1] from masked maps Wmap[B,L,H,W] and grid[HW,2].

    Per lane (LSQ_layer.py:110-116):  Yk = Wk*Y ; Z = Yk^T Yk + reg*I ;
    beta = Z^-1 Yk^T (Wk*x)  with  y = const - grid[:,1]  (:94; const=255 BP, 1 BEV).
    ``skip_rows`` > 0 drops the first rows from the sums instead of multiplying them
    by zero (the 0*inf = NaN convention of SURVEY.md 7.2 #10); with finite grids the
    two are identical.
    Also returns Z^-1 (float of the same dtype) for the closed-form backward.
    """
    B, L, H, Wd = Wmap.shape
    dt = Wmap.dtype
    grid = grid.to(dt)
    x = grid[:, 0]
    y = (const - grid[:, 1])
    Wf = Wmap.reshape(B, L, H * Wd)
    if skip_rows > 0:
        x, y, Wf = x[skip_rows * Wd:], y[skip_rows * Wd:], Wf[:, :, skip_rows * Wd:]
    Y = design_matrix(y, order)                      # [P, d+1]
    Yk = Wf.unsqueeze(-1) * Y                        # [B,L,P,d+1]
    Z = Yk.transpose(-1, -2) @ Yk + reg_ls * torch.eye(order + 1, dtype=dt)
    X = Yk.transpose(-1, -2) @ (Wf * x).unsqueeze(-1)
    if use_cholesky:                                 # Networks/gels.py:11-15
        U = torch.linalg.cholesky(Z, upper=True)
        beta = torch.cholesky_solve(X, U, upper=True)
        Zinv = torch.cholesky_inverse(U, upper=True)
    else:
        Zinv = torch.inverse(Z)                      # LSQ_layer.py:114
        beta = Zinv @ X                              # :116
    return beta.squeeze(-1), Zinv


def wls_backward_closed_form(o, grid, order, beta, Zinv, gbeta, act="square",
                             zero_rows=0, const=255.0):
    """d loss / d o  for  o -> activation -> mask -> wls_forward  (float64 numpy).

    SURVEY.md Appendix C / Networks/gels.py:18-25:  z = Z^-1 g ;
    dL/d(mw)_p = 2 (mw)_p (x_p - phi_p^T beta)(phi_p^T z) ;  dL/do = m a'(o) dL/d(mw).
    """
    o = np.asarray(o, dtype=np.float64)
    B, L, H, Wd = o.shape
    g = np.asarray(grid, dtype=np.float64)
    x = g[:, 0].reshape(H, Wd)
    y = (const - g[:, 1]).reshape(H, Wd)
    ot = torch.from_numpy(o).requires_grad_(True)
    a = activation(ot, act)
    (da,) = torch.autograd.grad(a.sum(), ot)
    a = a.detach().numpy()
    da = da.numpy()
    m = np.ones((H, 1))
    m[:zero_rows] = 0
    z = np.einsum("blij,blj->bli", np.asarray(Zinv, np.float64), np.asarray(gbeta, np.float64))
    pw = np.stack([y ** k for k in range(order, -1, -1)], 0)        # [d+1,H,W]
    fit = np.einsum("bli,ihw->blhw", np.asarray(beta, np.float64), pw)
    zz = np.einsum("bli,ihw->blhw", z, pw)
    return m * da * 2.0 * (m * a) * (x - fit) * zz


# ---------------------------------------------------------------------------
# Losses on beta
# ---------------------------------------------------------------------------

class BackprojectionLoss:
    """Backprojection_Loss/Loss_crit.py:161-218 (float64)."""

    def __init__(self, order, resize=256, no_mapping=False, M=None, M_inv=None):
        if M is None:
            M, M_inv = get_homography(resize, no_mapping)
        self.M = torch.as_tensor(M, dtype=torch.float64)
        self.M_inv = torch.as_tensor(M_inv, dtype=torch.float64)
        start, delta = 160, 10
        self.y_d = (torch.arange(start, 720, delta) - 80).double() / 2.5          # :173
        self.y_prime = (self.M[1, 1:2] * self.y_d + self.M[1, 2:]) / \
                       (self.M[2, 1:2] * self.y_d + self.M[2, 2:])                  # :175
        self.y_eval = 255 - self.y_prime                                           # :176
        self.Y = design_matrix(self.y_eval, order)                                 # :178-188

    def __call__(self, params, x_gt, valid):
        """params [B,d+1,1] or [B,d+1]; x_gt, valid [B,56] -> (loss, x_cal*valid)."""
        p = params.reshape(params.shape[0], -1).double()
        x_prime = p @ self.Y.t()                                                   # :205
        ones = torch.ones_like(x_prime)
        coords = torch.stack((x_prime, self.y_prime.expand_as(x_prime), ones), 1)  # :208
        trans = self.M_inv.unsqueeze(0) @ coords                                   # :209
        x_cal = trans[:, 0, :] / trans[:, 2, :]                                    # :210
        x_err = (x_gt - x_cal) * valid                                             # :214
        loss = torch.sum(x_err ** 2) / valid.sum()                                 # :215
        if valid.sum() == 0:
            loss = 0
        return loss, x_cal * valid


def area_loss(params, gt_params, order=2, weight_funct="none"):
    """Backprojection_Loss/Loss_crit.py:98-143 == Birds_Eye_View_Loss/Loss_crit.py:98-134,
    with the torch>=2 incompatible ``.byte()`` mask (:140-141) restated as bool."""
    diff = params.reshape(params.shape[0], -1) - gt_params
    a, b = diff[:, 0], diff[:, 1]
    t = 0.7
    if order == 2:
        c = diff[:, 2]
        if weight_funct == "none":
            loss_fit = (a**2)*(t**5)/5 + 2*a*b*(t**4)/4 + (b**2 + c*2*a)*(t**3)/3 + 2*b*c*(t**2)/2 + (c**2)*t
        elif weight_funct == "linear":
            loss_fit = c**2*t - t**5*((2*a*b)/5 - a**2/5) + t**2*(b*c - c**2/2) - (a**2*t**6)/6 - \
                t**4*(b**2/4 - (a*b)/2 + (a*c)/2) + t**3*(b**2/3 - (2*c*b)/3 + (2*a*c)/3)
        elif weight_funct == "quadratic":
            loss_fit = t**3*(1/3*b**2 + 2/3*a*c) - t**(7/2)*(2/7*b**2 + 4/7*a*c) + c**2*t + 0.2*a**2*t**5 - \
                2/11*a**2*t**(11/2) - 2/3*c**2*t**(3/2) + 0.5*a*b*t**4 - 4/9*a*b*t**(9/2) + b*c*t**2 - \
                0.8*b*c*t**(5/2)
        else:
            raise NotImplementedError(weight_funct)
    elif order == 1:
        loss_fit = (b**2)*t + a*b*(t**2) + ((a**2)*(t**3))/3
    else:
        raise NotImplementedError(order)
    mask = torch.prod((gt_params != 0).to(torch.int64), 1).bool()
    sel = loss_fit[mask]
    return sel.mean(0) if sel.numel() != 0 else torch.zeros((), dtype=loss_fit.dtype)
