"""``GELS`` -- the Cholesky normal-equation solver of the reference
(BP/Networks/gels.py:9-25) as an autograd Function with the same call signature
``GELS.apply(A, b)``:  A [B,P,n] (= W*Y), b [B,P,1] (= W*x)  ->  x [B,n,1].

The hot path (``Weighted_least_squares`` with ``use_cholesky=True``) never builds A or
b: csrc/lsq.cu forms the moments directly from the map and runs the Cholesky solve and
this backward formula in-kernel (LF_SOLVER_CHOLESKY).  This class only keeps the
reference's public name alive for callers that hold explicit A, b; it is NOT on the
measured path and uses torch's device linalg for the n x n factorisation.
"""
import torch

from ._pkg import capi


class GELS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, b):
        capi().require_cuda(A, b)
        # normal equations in fp64 on device (tiny: n x n per batch element)
        Ad, bd = A.double(), b.double()
        Z = Ad.transpose(-1, -2) @ Ad
        U = torch.linalg.cholesky(Z, upper=True)
        x = torch.cholesky_solve(Ad.transpose(-1, -2) @ bd, U, upper=True).to(A.dtype)
        ctx.save_for_backward(U, x, A, b)
        return x

    @staticmethod
    def backward(ctx, grad_output):
        U, x, A, b = ctx.saved_tensors
        z = torch.cholesky_solve(grad_output.double(), U, upper=True).to(A.dtype)
        xz = x @ z.transpose(-1, -2)
        grad_A = -A @ (xz + xz.transpose(-1, -2)) + b @ z.transpose(-1, -2)
        grad_b = A @ z
        return grad_A, grad_b
