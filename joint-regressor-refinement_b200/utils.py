"""Call-compatible mirrors of the reference's loss / geometry helpers
(``scripts/utils.py:85-215``, ``scripts/eval_utils.py:7-58``).  ``find_joints`` is the one on
the hot path: without autograd it runs the fused CUDA kernels (vertices never materialised);
with autograd it goes through ``SMPLFunction`` so gradients reach betas / rotations / J."""
from __future__ import annotations

import random

import numpy as np
import torch

from ._lib import POSE_ROTMAT


def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    """scripts/utils.py:190-204 ([B,6] interleaved -> [B,3,3], columns b1,b2,b3)."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = torch.nn.functional.normalize(a1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def find_j_reg_mask(j_reg: torch.Tensor) -> torch.Tensor:
    """scripts/utils.py:182-187.  The reference builds the mask from two all-ones tensors, so
    it is all ones; reproduced (sparsity is preserved by relu'(x<=0)=0 instead)."""
    return torch.ones_like(j_reg)


class _RegressorInstalled:
    """The CUDA model keeps ONE normalised regressor, shared with PoseRefiner / RegressorRefit; the reference's find_joints
    is stateless (utils.py:87-92 re-normalises on every call).  When a call's J is not the one the model holds it is
    installed for the call and the previous one is put back afterwards, so a comparison run with another regressor does
    not change what the refinement loop optimises against."""

    def __init__(self, native, J, mask):
        self.native, self.J, self.mask, self.prev = native, J, mask, None

    def __enter__(self):
        if not self.native.holds_regressor(self.J, self.mask):
            self.prev = self.native._reg
            self.native.set_regressor(self.J, self.mask)

    def __exit__(self, *exc):
        if self.prev is not None:
            self.native.set_regressor(self.prev[0], self.prev[1])
        return False


class FindJointsFunction(torch.autograd.Function):
    """(betas [B,10], rotations [B,24,9]) -> joints17 [B,17,3] through the fused loss-path kernels; the backward
    (jrr_find_joints_backward) gives the gradients w.r.t. betas and the rotation matrices.  The regressor is a constant
    here (its own gradient is the refit's business: RegressorRefit / the eager path of find_joints)."""

    @staticmethod
    def forward(ctx, betas, full, native, J, mask):
        betas, full = betas.float().contiguous(), full.float().contiguous()
        with _RegressorInstalled(native, J, mask):
            out = native.find_joints(betas, full, POSE_ROTMAT)
        ctx.save_for_backward(betas, full)
        ctx.native, ctx.J, ctx.mask = native, J, mask
        return out

    @staticmethod
    def backward(ctx, dpred):
        betas, full = ctx.saved_tensors
        with _RegressorInstalled(ctx.native, ctx.J, ctx.mask):
            dbetas, dfull = ctx.native.find_joints_backward(betas, full, POSE_ROTMAT, dpred)
        return dbetas, dfull, None, None, None


def find_joints(smpl, shape, orient, pose, J_regressor, mask=None, return_verts=False):
    """scripts/utils.py:85-103: regressed 17 joints [B,17,3] (and optionally the vertices).  Differentiable w.r.t. the
    body-model inputs through the fused kernels; only when the REGRESSOR itself requires grad (or the vertices are asked
    for) does it compose the module forward with a torch contraction like the reference."""
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (shape, orient, pose, J_regressor))
    if return_verts or (needs_grad and J_regressor.requires_grad):
        J = J_regressor * mask if mask is not None else J_regressor
        Jn = torch.relu(J)
        Jn = Jn / Jn.sum(dim=1, keepdim=True)
        verts = smpl(global_orient=orient, body_pose=pose, betas=shape, pose2rot=False).vertices
        pred = torch.matmul(Jn.to(verts.device)[None], verts)
        return (pred, verts) if return_verts else pred
    native = smpl.native()
    J = J_regressor.detach().to(native.device)
    m = None if mask is None else mask.detach().to(native.device)
    B = max(shape.shape[0], pose.shape[0])
    full = torch.cat([orient.reshape(-1, 1, 3, 3).expand(B, -1, -1, -1),
                      pose.reshape(-1, 23, 3, 3).expand(B, -1, -1, -1)], dim=1).reshape(B, 24, 9)
    betas = shape if shape.shape[0] == B else shape.expand(B, -1)
    if needs_grad:
        return FindJointsFunction.apply(betas, full, native, J, m)
    with _RegressorInstalled(native, J, m):
        return native.find_joints(betas, full, POSE_ROTMAT)


def move_pelvis(j3ds: torch.Tensor) -> torch.Tensor:
    """scripts/utils.py:106-114."""
    return j3ds - j3ds[:, [0], :]


def batch_compute_similarity_transform_torch(S1, S2):
    """scripts/eval_utils.py:7-58 (metric only, plain torch): similarity-align S1 to S2."""
    transposed = False
    if S1.shape[0] != 3 and S1.shape[0] != 2:
        S1, S2 = S1.permute(0, 2, 1), S2.permute(0, 2, 1)
        transposed = True
    mu1, mu2 = S1.mean(dim=-1, keepdim=True), S2.mean(dim=-1, keepdim=True)
    X1, X2 = S1 - mu1, S2 - mu2
    var1 = (X1 ** 2).sum(dim=(1, 2))
    K = X1 @ X2.transpose(1, 2)
    U, _, Vh = torch.linalg.svd(K)
    Vm = Vh.transpose(1, 2)
    Z = torch.eye(U.shape[1], device=S1.device, dtype=S1.dtype).repeat(U.shape[0], 1, 1)
    Z[:, -1, -1] *= torch.sign(torch.det(U @ Vh))
    R = Vm @ Z @ U.transpose(1, 2)
    scale = torch.diagonal(R @ K, dim1=1, dim2=2).sum(1) / var1
    t = mu2 - scale[:, None, None] * (R @ mu1)
    out = scale[:, None, None] * (R @ S1) + t
    return out.permute(0, 2, 1) if transposed else out


def evaluate(pred_j3ds, target_j3ds, per_frame=False):
    """scripts/utils.py:117-145: (MPJPE, PA-MPJPE) in mm; target in mm, prediction in m.
    CUDA tensors run the hand-written metric kernel (`jrr_evaluate`: per-frame Procrustes with a
    Jacobi 3x3 SVD); CPU tensors use the plain torch restatement below."""
    if pred_j3ds.is_cuda and pred_j3ds.shape[1:] == (17, 3):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        B = pred_j3ds.shape[0]
        p = pred_j3ds.detach().float().contiguous()
        t = target_j3ds.detach().float().contiguous().to(p.device)
        out = torch.empty(2, device=p.device)
        pf = torch.empty(B, 2, device=p.device) if per_frame else None
        scratch = torch.empty((B + 127) // 128 * 2 + 2, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(L.jrr_evaluate(B, C.c_void_p(p.data_ptr()), C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()),
                                      C.c_void_p(pf.data_ptr()) if pf is not None else None,
                                      C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel() * 8),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "jrr_evaluate")
        o = out.cpu().numpy()
        return (o[0], o[1], pf) if per_frame else (o[0], o[1])
    with torch.no_grad():
        p = move_pelvis(pred_j3ds.detach().clone().float())
        t = move_pelvis(target_j3ds.detach().clone().float() / 1000)
        mpjpe = ((p - t) ** 2).sum(-1).sqrt().mean(-1).cpu().numpy().mean() * 1000
        pa = ((batch_compute_similarity_transform_torch(p, t) - t) ** 2).sum(-1).sqrt().mean(-1).cpu().numpy().mean() * 1000
    return mpjpe, pa


def set_seed(seed):
    """scripts/utils.py:207-215."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
