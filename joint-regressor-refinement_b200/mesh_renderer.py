"""Drop-in for the reference's silhouette renderer, ``scripts/mesh_renderer.py:23-79`` (pytorch3d 0.3.0
``MeshRasterizer(blur_radius=0, faces_per_pixel=1)`` + ``SoftSilhouetteShader(sigma=1e-4)`` behind
``PerspectiveCameras(T=batch['cam'], focal_length=5000/image_size)``) and for ``render_mesh``
(``scripts/optimize.py:77-85``), on hand-written sm_100a kernels (``csrc/jrr_silhouette.cu``) behind
``torch.autograd.Function`` -- differentiable w.r.t. the mesh vertices and the camera translation, so the
silhouette term of ``optimize.py:234-236,252-253`` composes with ``SMPL`` exactly as in the reference:

    renderer = Mesh_Renderer(image_size=224, faces=smpl.faces)
    img = render_mesh(smpl, renderer, betas, orient, pose, batch)          # [B,1,224,224]
    loss = torch.nn.functional.mse_loss(img, batch["mask_rcnn"]) * 100

``silhouette_mse`` is the fused form of the last two lines (loss and gradient seed inside the kernels).
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import JrrError, check
from .native import _f32c, _ptr, _stream

SIGMA = 1e-4        # BlendParams(sigma=1e-4, gamma=1e-4), mesh_renderer.py:29


def load_obj_faces(path: str) -> np.ndarray:
    """Vertex indices of the ``f`` records of a Wavefront OBJ (what ``load_obj(...)[1].verts_idx`` holds,
    mesh_renderer.py:41-42); polygons are fanned into triangles."""
    faces = []
    with open(path) as fh:
        for line in fh:
            if not line.startswith("f "):
                continue
            idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
            for k in range(1, len(idx) - 1):
                faces.append((idx[0], idx[k], idx[k + 1]))
    return np.asarray(faces, dtype=np.int64)


def vertex_face_csr(faces, n_verts: int):
    """CSR vertex -> (face * 3 + corner) entries, ascending inside a vertex's list (the fixed order in which the backward
    sums a vertex's corner gradients).  Returns (faces int32 [F,3], ptr int32 [V+1], idx int32 [3F])."""
    f = np.ascontiguousarray(np.asarray(faces).reshape(-1, 3), dtype=np.int64)
    if f.size == 0 or f.min() < 0 or f.max() >= n_verts:
        raise JrrError("faces must index the mesh vertices")
    flat = f.reshape(-1)
    order = np.argsort(flat, kind="stable")
    ptr = np.zeros(n_verts + 1, dtype=np.int64)
    np.add.at(ptr, flat + 1, 1)
    return f.astype(np.int32), np.cumsum(ptr).astype(np.int32), order.astype(np.int32)


class _Mesh:
    """Device copies of the faces and the vertex -> (face, corner) CSR the backward gathers through."""

    def __init__(self, faces, n_verts: int, device):
        f, ptr, idx = vertex_face_csr(faces, n_verts)
        self.F, self.V = int(f.shape[0]), int(n_verts)
        self.faces = torch.from_numpy(f).to(device)
        self.vf_ptr = torch.from_numpy(ptr).to(device)
        self.vf_idx = torch.from_numpy(idx).to(device)
        self._ws = None

    def workspace(self, B: int, S: int):
        need = _lib.lib().jrr_silhouette_workspace_bytes(B, self.V, self.F, S)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.faces.device)
        return C.c_void_p(self._ws.data_ptr()), C.c_size_t(self._ws.numel())


def _forward(mesh: _Mesh, verts, cam, S, flip_scale, target=None, logical_batch=0):
    L = _lib.lib()
    B = verts.shape[0]
    alpha = torch.empty(B, S, S, device=verts.device)
    p2f = torch.empty(B, S, S, dtype=torch.int32, device=verts.device)
    loss = torch.empty(1, device=verts.device) if target is not None else None
    ws, wsz = mesh.workspace(B, S)
    with torch.cuda.device(verts.device):
        check(L.jrr_silhouette_forward(B, _ptr(verts), mesh.V, _ptr(cam), _ptr(mesh.faces), mesh.F, S, 5000.0 / S, SIGMA,
                                       int(flip_scale), _ptr(target), logical_batch, _ptr(alpha), _ptr(p2f), _ptr(loss),
                                       ws, wsz, _stream()), "jrr_silhouette_forward")
    return alpha, p2f, loss


def _backward(mesh: _Mesh, verts, cam, S, flip_scale, alpha, p2f, dalpha=None, target=None, logical_batch=0, weight=1.0):
    L = _lib.lib()
    B = verts.shape[0]
    dverts = torch.empty_like(verts)
    dcam = torch.empty(B, 3, device=verts.device)
    ws, wsz = mesh.workspace(B, S)
    with torch.cuda.device(verts.device):
        check(L.jrr_silhouette_backward(B, _ptr(verts), mesh.V, _ptr(cam), _ptr(mesh.faces), mesh.F, _ptr(mesh.vf_ptr),
                                        _ptr(mesh.vf_idx), S, 5000.0 / S, SIGMA, int(flip_scale), _ptr(alpha), _ptr(p2f),
                                        _ptr(dalpha), _ptr(target), logical_batch, float(weight), _ptr(dverts), _ptr(dcam),
                                        ws, wsz, _stream()), "jrr_silhouette_backward")
    return dverts, dcam


class SilhouetteFunction(torch.autograd.Function):
    """(vertices [B,V,3], cam [B,3]) -> alpha [B,S,S].  The backward re-projects (the workspace may have been reused by
    another call in between) and then runs the face / vertex / camera gradient kernels."""

    @staticmethod
    def forward(ctx, verts, cam, mesh, S, flip_scale):
        verts, cam = _f32c(verts, "vertices"), _f32c(cam, "cam")
        alpha, p2f, _ = _forward(mesh, verts, cam, S, flip_scale)
        ctx.save_for_backward(verts, cam, alpha, p2f)
        ctx.mesh, ctx.S, ctx.flip_scale = mesh, S, flip_scale
        ctx.mark_non_differentiable(p2f)
        return alpha, p2f

    @staticmethod
    def backward(ctx, dalpha, _dp2f):
        verts, cam, alpha, p2f = ctx.saved_tensors
        _forward(ctx.mesh, verts, cam, ctx.S, ctx.flip_scale)          # projected vertices back into the workspace
        dverts, dcam = _backward(ctx.mesh, verts, cam, ctx.S, ctx.flip_scale, alpha, p2f,
                                 dalpha=_f32c(dalpha, "dalpha"))
        return dverts, dcam, None, None, None


class Mesh_Renderer(nn.Module):
    """``Mesh_Renderer(image_size)(batch, mesh_verts) -> [B,4,S,S]`` (mesh_renderer.py:23-79): channels 0-2 are the
    constant white texture, channel 3 the soft silhouette.  ``faces``: the mesh topology ([F,3]; the reference reads it
    from data/body_model/smpl_uv.obj -- pass ``obj_path`` for that)."""

    def __init__(self, image_size=256, faces=None, obj_path=None):
        super().__init__()
        if faces is None:
            if obj_path is None:
                raise ValueError("give faces ([F,3]) or obj_path")
            faces = load_obj_faces(obj_path)
        self.image_size = int(image_size)
        self.faces = np.asarray(faces).reshape(-1, 3)
        self._mesh = None

    def mesh(self, n_verts: int, device) -> _Mesh:
        if device.type != "cuda":
            raise JrrError("the silhouette renderer needs CUDA tensors (no CPU path)")
        m = self._mesh
        if m is None or m.V != n_verts or m.faces.device != device:
            self._mesh = m = _Mesh(self.faces, n_verts, device)
        return m

    def silhouette(self, cam, verts, flip_scale=False):
        alpha, _ = SilhouetteFunction.apply(verts, cam, self.mesh(verts.shape[1], verts.device), self.image_size, flip_scale)
        return alpha

    def forward(self, batch, smpl_verts):
        alpha = self.silhouette(batch["cam"], smpl_verts, False)
        ones = torch.ones_like(alpha)
        return torch.stack([ones, ones, ones, alpha], dim=1)


def render_mesh(smpl, silhouette_renderer, betas, orient, pose, batch):
    """optimize.py:77-85.  (The flip / scale of the vertices is applied inside the projection kernel.)"""
    verts = smpl(global_orient=orient, body_pose=pose, betas=betas, pose2rot=False).vertices
    return silhouette_renderer.silhouette(batch["cam"], verts, True).unsqueeze(1)


def silhouette_mse(silhouette_renderer, verts, cam, target, logical_batch=None, weight=1.0):
    """Fused ``weight * MSELoss(render, target)`` on the body model's vertices (optimize.py:234-236 with the weight of
    optimize.py:252): returns (loss / weight as a 1-element tensor, d/d vertices, d/d cam, alpha) in two calls, with the
    loss reduction and the gradient seed inside the kernels."""
    verts, cam = _f32c(verts, "vertices"), _f32c(cam, "cam")
    S = silhouette_renderer.image_size
    target = _f32c(target, "target").reshape(verts.shape[0], S, S)
    mesh = silhouette_renderer.mesh(verts.shape[1], verts.device)
    LB = verts.shape[0] if logical_batch is None else int(logical_batch)
    alpha, p2f, loss = _forward(mesh, verts, cam, S, True, target, LB)
    dverts, dcam = _backward(mesh, verts, cam, S, True, alpha, p2f, target=target, logical_batch=LB, weight=weight)
    return loss, dverts, dcam, alpha
