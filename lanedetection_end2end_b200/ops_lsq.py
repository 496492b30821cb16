"""Host side of the fused weighted least-squares layer: grid tables, workspace
and the ``torch.autograd.Function`` that calls ``lf_lsq_fwd`` / ``lf_lsq_bwd``
(include/lanefit_b200.h).  Torch is plumbing here (device memory, streams,
autograd graph); all arithmetic on the maps happens in csrc/lsq.cu.
"""
import torch

from . import _capi

_TABLE_CACHE = {}
_WS_CACHE = {}


class GridTables:
    """De-interleaved, batch-invariant view of the reference's ``grid`` [N, H*W, 2]
    (BP/Networks/LSQ_layer.py:50-68):  xtab = grid[0,:,0],  ytab = const - grid[0,:,1]
    (:93-94, fp32 like the reference), and yrow [H] when ytab is constant along rows
    (true for every homography ``get_homography`` builds) -> row-separable fast kernel.
    """

    def __init__(self, grid, H, W, const):
        g = grid[0] if grid.dim() == 3 else grid
        if g.shape[0] != H * W or g.shape[1] != 2:
            raise ValueError("grid must be [N, H*W, 2]; got %s for H=%d W=%d" % (tuple(grid.shape), H, W))
        g = g.detach().to(torch.float32)
        self.xtab = g[:, 0].contiguous()
        self.ytab = (const - g[:, 1]).contiguous()
        y2 = self.ytab.view(H, W)
        # Row-separable?  Mathematically y' depends on the row only for every homography the reference
        # builds (M[1,0], M[2,0] are round-off zeros), but its fp32 bmm leaves 1-ulp differences inside a
        # few rows (10 of 256 at --resize 256).  Rows whose spread is <= 4 ulp are snapped to their median
        # (effect on beta ~1e-7 relative, far below the reference's own fp32 noise).
        # Non-finite rows (the vanishing line at --resize 320) count as constant: they
        # must be masked anyway.
        finite = torch.isfinite(y2)
        ysafe = torch.where(finite, y2, torch.zeros_like(y2))
        spread = ysafe.max(dim=1).values - ysafe.min(dim=1).values
        # ulp of the quantity the reference rounded: max(|y'|, |const - y'|) per row
        yp = torch.where(finite, g[:, 1].view(H, W), torch.zeros_like(y2))
        scale = torch.maximum(ysafe.abs().max(dim=1).values, yp.abs().max(dim=1).values).clamp_min(1e-30)
        row_ok = (spread <= 4 * 1.1920929e-07 * scale) | (~finite).any(dim=1)
        self.rowsep = bool(row_ok.all().item()) and (W % 4 == 0)
        self.yrow = y2.median(dim=1).values.contiguous() if self.rowsep else None
        self.H, self.W, self.const = H, W, const


def grid_tables(grid, H, W, const):
    key = (grid.data_ptr(), grid._version, tuple(grid.shape), str(grid.device), H, W, float(const))
    t = _TABLE_CACHE.get(key)
    if t is None:
        if len(_TABLE_CACHE) > 16:
            _TABLE_CACHE.clear()
        t = GridTables(grid, H, W, const)
        _TABLE_CACHE[key] = t
    return t


def _workspace(device, nbytes):
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


def status_message(st):
    parts = []
    if st & _capi.STATUS_SINGULAR:
        parts.append("singular normal matrix (torch.inverse would raise)")
    if st & _capi.STATUS_NOT_POSDEF:
        parts.append("normal matrix not positive definite (torch.cholesky would raise)")
    if st & _capi.STATUS_NONFINITE:
        parts.append("non-finite moments or coefficients")
    return "weighted least squares: " + "; ".join(parts)


class LsqFunction(torch.autograd.Function):
    """beta, masked = LSQ(o).  o: [B,L,H,W] f32/bf16 raw (or pre-activated) maps."""

    @staticmethod
    def forward(ctx, o, tables, order, mask_rows, act, reg_ls, solver, want_masked, status_out):
        _capi.require_cuda(o)
        if order < 0 or order > _capi.MAX_ORDER:
            raise NotImplementedError("Requested order {} for polynomial fit is not implemented".format(order))
        o = o.contiguous()
        B, L, H, W = o.shape
        assert (H, W) == (tables.H, tables.W)
        n = order + 1
        h = _capi.lib()
        dev = o.device
        beta = torch.empty(B, L, n, dtype=torch.float64, device=dev)
        zinv = torch.empty(B, L, n, n, dtype=torch.float64, device=dev)
        masked = torch.empty(B, L, H, W, dtype=torch.float32, device=dev) if want_masked else None
        status = status_out if status_out is not None else torch.zeros(1, dtype=torch.int32, device=dev)
        nbytes = h.lf_lsq_workspace_bytes(B, L, H, W, order)
        ws = _workspace(dev, nbytes)
        esz = o.element_size()
        with torch.cuda.device(dev):
            _capi.call("lf_lsq_fwd", _capi.ptr(o), _capi.dtype_id(o), _capi.ptr(tables.xtab), _capi.ptr(tables.ytab),
                       _capi.ptr(tables.yrow), B, L, H, W, order, mask_rows, act, float(reg_ls), solver,
                       _capi.ptr(beta), _capi.ptr(zinv), _capi.ptr(masked), _capi.ptr(status),
                       _capi.ptr(ws), ws.numel(), _capi.stream_ptr(),
                       nbytes=B * L * H * W * (esz + (4 if want_masked else 0)))
        if status_out is None:
            # same implicit sync + RuntimeError as torch.inverse in the reference
            # (BP/Networks/LSQ_layer.py:114, caught at BP/main.py:289-292)
            st = int(status.item())
            if st:
                raise RuntimeError(status_message(st))
        ctx.save_for_backward(o, beta, zinv)
        ctx.tables = tables
        ctx.cfg = (order, mask_rows, act)
        if masked is not None:
            ctx.mark_non_differentiable(masked)
            return beta, masked
        return beta, None

    @staticmethod
    def backward(ctx, gbeta, _gmasked):
        o, beta, zinv = ctx.saved_tensors
        order, mask_rows, act = ctx.cfg
        t = ctx.tables
        B, L, H, W = o.shape
        gbeta = gbeta.to(torch.float64).contiguous()
        d_o = torch.empty_like(o)
        with torch.cuda.device(o.device):
            _capi.call("lf_lsq_bwd", _capi.ptr(o), _capi.dtype_id(o), _capi.ptr(t.xtab), _capi.ptr(t.ytab),
                       _capi.ptr(t.yrow), B, L, H, W, order, mask_rows, act, _capi.ptr(beta), _capi.ptr(zinv),
                       _capi.ptr(gbeta), _capi.ptr(d_o), _capi.stream_ptr(),
                       nbytes=2 * B * L * H * W * o.element_size())
        return d_o, None, None, None, None, None, None, None, None


def lsq(o, grid, order, const=255.0, mask_rows=0, act="none", reg_ls=0.0, use_cholesky=False,
        want_masked=False, status_out=None):
    """Functional entry: returns (beta [B,L,order+1] f64, masked or None)."""
    B, L, H, W = o.shape
    tables = grid_tables(grid, H, W, const)
    if act not in _capi.ACT_IDS:
        raise NotImplementedError("Activation type: {} is not implemented".format(act))
    solver = _capi.SOLVER_CHOLESKY if use_cholesky else _capi.SOLVER_INVERSE
    return LsqFunction.apply(o, tables, order, mask_rows, _capi.ACT_IDS[act], reg_ls, solver, want_masked,
                             status_out)
