"""Drop-in for the SO(3) part of the reference's ``util.py`` (same names, signatures, shapes).

Every function is ONE fused sm_100a kernel on CUDA float32 tensors (via the C ABI in
include/so3d.h); gradients are provided by custom autograd functions whose backward passes are
kernels too.  There is no CPU path.  Reference citations are file:line of
qazwsxal/diffusion-extensions @ f100885d.

Deliberate deviations from the reference:
  * exp maps use the Rodrigues closed form instead of matrix_exp (+ SVD), so results stay
    orthonormal for any scalar (Q5);
  * log_rmat / rmat_to_aa are accurate up to and at a rotation by pi (Q4) and return the axis
    (0,0,1) with angle 0 at the identity instead of NaN (Q10).
"""
from typing import Tuple

import torch

from . import ops


# ---------------------------------------------------------------------------------------------
# hat / vee  (util.py:79-92) -- pure data movement, kept as torch indexing (fused away inside the
# kernels wherever they sit on the hot path)
# ---------------------------------------------------------------------------------------------
def skew2vec(skew: torch.Tensor) -> torch.Tensor:
    return torch.stack((skew[..., 2, 1], -skew[..., 2, 0], skew[..., 1, 0]), dim=-1)


def vec2skew(vec: torch.Tensor) -> torch.Tensor:
    x, y, z = vec[..., 0], vec[..., 1], vec[..., 2]
    o = torch.zeros_like(x)
    return torch.stack((o, -z, y, z, o, -x, -y, x, o), dim=-1).reshape(vec.shape[:-1] + (3, 3))


def orthogonalise(mat):
    """util.py:95-107: U round(S) V^T of the 3x3 block.  Not on the hot path any more (the exp
    maps below are orthonormal by construction); kept for API compatibility."""
    orth_mat = mat.clone()
    u, s, vh = torch.linalg.svd(mat[..., :3, :3])
    orth_mat[..., :3, :3] = u @ torch.diag_embed(s.round()) @ vh
    return orth_mat


# ---------------------------------------------------------------------------------------------
# autograd functions
# ---------------------------------------------------------------------------------------------
class _LogRmat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r_mat):
        ctx.save_for_backward(r_mat)
        return ops.log_rmat(r_mat)

    @staticmethod
    def backward(ctx, grad):
        (r_mat,) = ctx.saved_tensors
        return ops.log_rmat_bwd(r_mat, grad.contiguous())


class _AaToRmat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axis, ang):
        # axis (...,3), ang (...,1) already broadcast to a common batch shape
        ctx.save_for_backward(axis, ang)
        return ops.aa_to_rmat(axis, ang[..., 0])

    @staticmethod
    def backward(ctx, grad):
        axis, ang = ctx.saved_tensors
        g_axis, g_ang = ops.aa_to_rmat_bwd(axis.contiguous(), ang[..., 0].contiguous(), grad.contiguous())
        return g_axis, g_ang[..., None]


class _ExpVec(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec):
        ctx.save_for_backward(vec)
        return ops.exp_vec(vec)

    @staticmethod
    def backward(ctx, grad):
        (vec,) = ctx.saved_tensors
        return ops.exp_vec_bwd(vec.contiguous(), grad.contiguous())


class _So3Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rmat, scalars):
        ctx.save_for_backward(rmat, scalars)
        return ops.so3_scale(rmat, scalars)

    @staticmethod
    def backward(ctx, grad):
        rmat, scalars = ctx.saved_tensors
        bs = torch.broadcast_shapes(rmat.shape[:-2], scalars.shape) if scalars.numel() > 1 else rmat.shape[:-2]
        r = rmat.expand(*bs, 3, 3).contiguous()
        s, stride = ops._per_row(scalars, bs, r.device, "scalars")
        g_r, g_s = ops.so3_scale_bwd(r, s, stride, grad.expand(*bs, 3, 3).contiguous())
        g_r = g_r.sum_to_size(rmat.shape) if g_r.shape != rmat.shape else g_r
        g_s = g_s.sum_to_size(scalars.shape) if g_s.shape != scalars.shape else g_s
        return (g_r if ctx.needs_input_grad[0] else None), (g_s if ctx.needs_input_grad[1] else None)


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


# ---------------------------------------------------------------------------------------------
# public API (reference names)
# ---------------------------------------------------------------------------------------------
def log_rmat(r_mat: torch.Tensor) -> torch.Tensor:
    """util.py:164-192: matrix logarithm of rotation matrices, (...,3,3) skew-symmetric."""
    if _needs_grad(r_mat):
        return _LogRmat.apply(r_mat)
    return ops.log_rmat(r_mat)


def log_vec(r_mat: torch.Tensor) -> torch.Tensor:
    """skew2vec(log_rmat(R)) in one kernel (the composition used at diffusion.py:355)."""
    if _needs_grad(r_mat):
        return skew2vec(_LogRmat.apply(r_mat))
    return ops.log_vec(r_mat)


def aa_to_rmat(rot_axis: torch.Tensor, ang: torch.Tensor):
    """util.py:195-205: rotation matrix from axis (...,3) (normalised inside) and angle (...,1)."""
    if ang.shape[-1:] != (1,):
        raise ValueError("ang must have a trailing dimension of 1, like the reference (ang[..., None] is applied to it)")
    bs = torch.broadcast_shapes(rot_axis.shape[:-1], ang.shape[:-1])
    if _needs_grad(rot_axis, ang):
        return _AaToRmat.apply(rot_axis.expand(*bs, 3), ang.expand(*bs, 1))
    return ops.aa_to_rmat(rot_axis, ang[..., 0])


def exp_vec(vec: torch.Tensor) -> torch.Tensor:
    """matrix_exp(vec2skew(vec)) (diffusion.py:294) as one kernel."""
    if _needs_grad(vec):
        return _ExpVec.apply(vec)
    return ops.exp_vec(vec)


def rmat_to_aa(r_mat) -> Tuple[torch.Tensor, torch.Tensor]:
    """util.py:208-219: (axis (...,3), angle (...,1)), angle in [0, pi]."""
    if _needs_grad(r_mat):
        skew_vec = skew2vec(_LogRmat.apply(r_mat))
        angle = skew_vec.norm(p=2, dim=-1, keepdim=True)
        return skew_vec / angle, angle
    axis, angle = ops.rmat_to_aa(r_mat)
    return axis, angle[..., None]


def quat_to_rmat(quaternions: torch.Tensor) -> torch.Tensor:
    """util.py:222-252: real-first quaternions (...,4) -> (...,3,3)."""
    if _needs_grad(quaternions):
        r, i, j, k = torch.unbind(quaternions, -1)
        two_s = 2.0 / (quaternions * quaternions).sum(-1)
        o = torch.stack(
            (
                1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
            ),
            -1,
        )
        return o.reshape(quaternions.shape[:-1] + (3, 3))
    return ops.quat_to_rmat(quaternions)


def rmat_to_quat(r_mat: torch.Tensor) -> torch.Tensor:
    """New: unit quaternion (real-first, real part >= 0); inverse of quat_to_rmat."""
    return ops.rmat_to_quat(r_mat)


def rmat_dist(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """util.py:315-322: geodesic distance as the Frobenius norm of log(input^T target)."""
    if _needs_grad(input, target):
        mul = input.transpose(-1, -2) @ target
        return _LogRmat.apply(mul).norm(p=2, dim=(-1, -2))
    return ops.rmat_dist(input, target)


def so3_lerp(rot_a: torch.Tensor, rot_b: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """util.py:325-338: weight has a trailing dimension of 1 (it multiplies the (...,1) angle)."""
    if _needs_grad(rot_a, rot_b, weight):
        rot_c = rot_a.transpose(-1, -2) @ rot_b
        axis, angle = rmat_to_aa(rot_c)
        return rot_a @ aa_to_rmat(axis, weight * angle)
    w = weight[..., 0] if isinstance(weight, torch.Tensor) and weight.dim() > 0 and weight.shape[-1] == 1 else weight
    return ops.so3_lerp(rot_a, rot_b, w)


def so3_scale(rmat, scalars):
    """util.py:349-361: scale the rotation angle, exp(scalars * log(rmat))."""
    if not isinstance(scalars, torch.Tensor):
        scalars = torch.tensor(float(scalars), dtype=torch.float32, device=rmat.device)
    if _needs_grad(rmat, scalars):
        return _So3Scale.apply(rmat, scalars)
    return ops.so3_scale(rmat, scalars)


def compose(a, b, trans_a=False, trans_b=False):
    """Batched 3x3 product op(a) @ op(b) as one kernel (no autograd; use `@` when gradients are needed)."""
    return ops.compose(a, b, trans_a, trans_b)


__all__ = [
    "skew2vec", "vec2skew", "orthogonalise", "log_rmat", "log_vec", "aa_to_rmat", "exp_vec", "rmat_to_aa",
    "quat_to_rmat", "rmat_to_quat", "rmat_dist", "so3_lerp", "so3_scale", "compose",
]
