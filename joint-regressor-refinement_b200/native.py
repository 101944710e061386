"""Thin object wrapper over the C ABI (include/jrr.h): owns the JrrModel handle and the
caller-side device workspace.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import POSE_AXIS_ANGLE, POSE_ROT6D, POSE_ROTMAT, JrrError, check  # noqa: F401


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    # the raw handle of torch's current stream on the current device (no Stream object: ~3 us less per call, which a
    # single-pose forward notices)
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise JrrError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def flatten_critic_state_dict(sd: dict) -> torch.Tensor:
    """state_dict of scripts/discriminator.py:Discriminator -> the flat parameter vector
    jrr_critic_load expects (registration order: conv_operations, linears, linear_operations)."""
    keys = ["conv_operations.0.weight", "conv_operations.0.bias",
            "conv_operations.2.weight", "conv_operations.2.bias"]
    for i in range(24):
        keys += [f"linears.{i}.weight", f"linears.{i}.bias"]
    for i in (0, 2, 4):
        keys += [f"linear_operations.{i}.weight", f"linear_operations.{i}.bias"]
    flat = torch.cat([sd[k].detach().float().reshape(-1) for k in keys])
    if flat.numel() != _lib.CRITIC_PARAMS:
        raise JrrError(f"critic state_dict has {flat.numel()} parameters, expected {_lib.CRITIC_PARAMS}")
    return flat


def flatten_shape_critic_state_dict(sd: dict) -> torch.Tensor:
    """state_dict of scripts/discriminator.py:Shape_Discriminator -> the 171 floats
    jrr_shape_critic_load expects."""
    keys = [f"shape_operations.{i}.{n}" for i in (0, 2, 4) for n in ("weight", "bias")]
    flat = torch.cat([sd[k].detach().float().reshape(-1) for k in keys])
    if flat.numel() != 171:
        raise JrrError(f"shape critic state_dict has {flat.numel()} parameters, expected 171")
    return flat


CRITIC_KEYS = (["conv_operations.0.weight", "conv_operations.0.bias", "conv_operations.2.weight",
                "conv_operations.2.bias"]
               + [f"linears.{i}.{n}" for i in range(24) for n in ("weight", "bias")]
               + [f"linear_operations.{i}.{n}" for i in (0, 2, 4) for n in ("weight", "bias")])
CRITIC_SHAPES = ([(32, 6, 1, 1), (32,), (32, 32, 1, 1), (32,)] + [(1, 32), (1,)] * 24
                 + [(1024, 768), (1024,), (1024, 1024), (1024,), (1, 1024), (1,)])
SHAPE_CRITIC_KEYS = [f"shape_operations.{i}.{n}" for i in (0, 2, 4) for n in ("weight", "bias")]
SHAPE_CRITIC_SHAPES = [(10, 10), (10,), (5, 10), (5,), (1, 5), (1,)]


def unflatten_state_dict(flat: torch.Tensor, keys, shapes) -> dict:
    """Inverse of flatten_*_state_dict: flat parameter vector -> state_dict with the reference's keys."""
    out, o = {}, 0
    for k, shp in zip(keys, shapes):
        n = 1
        for d in shp:
            n *= d
        out[k] = flat[o:o + n].reshape(shp).clone()
        o += n
    assert o == flat.numel()
    return out


class NativeModel:
    """Device-resident packed body model + regressor + critic (JrrModel*)."""

    def __init__(self, model: dict, device: torch.device, gemm_impl: int = 0):
        self.L = _lib.lib()
        if not torch.cuda.is_available():
            raise JrrError("CUDA device required: the hot path has no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise JrrError("NativeModel needs a CUDA device")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        f = lambda k: np.ascontiguousarray(np.asarray(model[k], dtype=np.float32))
        i = lambda k: np.ascontiguousarray(np.asarray(model[k], dtype=np.int64))
        keep = {k: f(k) for k in ("v_template", "shapedirs", "posedirs", "J_regressor",
                                  "lbs_weights", "J_regressor_extra")}
        keep.update({k: i(k) for k in ("parents", "joint_map", "vertex_picks")})
        assert keep["v_template"].shape == (6890, 3) and keep["shapedirs"].shape == (6890, 3, 10)
        assert keep["posedirs"].shape == (207, 20670) and keep["J_regressor"].shape == (24, 6890)
        assert keep["lbs_weights"].shape == (6890, 24) and keep["J_regressor_extra"].shape == (9, 6890)
        d = _lib.JrrModelDesc()
        for k, a in keep.items():
            setattr(d, k + "_host", a.ctypes.data)
        d.device = idx
        d.gemm_impl = gemm_impl
        h = C.c_void_p()
        check(self.L.jrr_model_create(C.byref(d), C.byref(h)), "jrr_model_create")
        self.h = h
        self._ws = None
        self._ws_B = 0
        self.launches = 0
        # generation counters captured CUDA graphs are keyed on (refine.py): the workspace pointer, the
        # regressor / packing, and everything else a captured launch sequence bakes in (loss path, shape critic)
        self.ws_generation = 0
        self.regressor_version = 0
        self.config_generation = 0
        # who set the current regressor: (J tensor, mask tensor, J._version) -- holders re-assert theirs
        # (PoseRefiner / RegressorRefit) and utils.find_joints restores it after a call with another J
        self._reg = None

    def _on_device(self):
        """Context that makes this model's device current -- a no-op (no context switch, ~10 us saved per call, which
        matters for single-pose forwards) when it already is."""
        if torch.cuda.current_device() == self.device.index:
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.jrr_model_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ workspace
    def workspace(self, B: int):
        if self._ws is None or self._ws_B < B:
            need = self.L.jrr_workspace_bytes(self.h, B)
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            self._ws_B = B
            self.ws_generation += 1      # graphs that captured the old pointer must be re-captured
        off = (-self._ws.data_ptr()) % 256
        return C.c_void_p(self._ws.data_ptr() + off), C.c_size_t(self._ws.numel() - off)

    def _done(self):
        self.launches = int(self.L.jrr_last_launch_count())

    # ------------------------------------------------------------------ state
    def set_regressor(self, J17_raw: torch.Tensor, mask: torch.Tensor | None = None):
        J = _f32c(J17_raw.detach(), "J_regressor")
        if tuple(J.shape) != (17, 6890):
            raise JrrError(f"J_regressor must be [17,6890], got {tuple(J.shape)}")
        m = _f32c(mask.detach(), "mask") if mask is not None else None
        with self._on_device():
            check(self.L.jrr_set_regressor(self.h, _ptr(J), _ptr(m), _stream()), "jrr_set_regressor")
        # the packed vertex order (and with it the workspace layout) may have been rebuilt
        self._ws = None
        self._ws_B = 0
        self.regressor_version += 1
        self._reg = (J17_raw, mask, J17_raw._version)

    def holds_regressor(self, J17_raw, mask=None) -> bool:
        """True when the model's normalised regressor was built from exactly this tensor (same storage, not
        modified through torch since) and mask."""
        r = self._reg
        return (r is not None and r[0].data_ptr() == J17_raw.data_ptr() and r[0].device == J17_raw.device
                and r[2] == J17_raw._version
                and ((r[1] is None) == (mask is None)) and (mask is None or r[1].data_ptr() == mask.data_ptr()))

    def load_critic(self, state_dict: dict):
        flat = flatten_critic_state_dict(state_dict).to(self.device)
        with self._on_device():
            check(self.L.jrr_critic_load(self.h, _ptr(flat), _stream()), "jrr_critic_load")
            torch.cuda.current_stream().synchronize()   # `flat` may be freed after return

    def critic_layer2_bwd_products(self, B: int) -> int:
        """Tensor-core products per K step of the critic's layer-2 backward GEMM in a refinement step (2 or 3)."""
        return int(self.L.jrr_critic_layer2_bwd_products(self.h, int(B)))

    def set_loss_path(self, mode: str):
        """'vertex' (per-vertex fused kernels, default) or 'folded' (regressor o skinning o blend operator
        folded once per regressor version; see include/jrr.h)."""
        code = {"vertex": 0, "folded": 1}[mode]
        with self._on_device():
            check(self.L.jrr_set_loss_path(self.h, code, _stream()), "jrr_set_loss_path")
        self.loss_path = mode
        self.config_generation += 1     # captured graphs hold the old launch sequence

    def load_shape_critic(self, state_dict, w_shape: float = 10.0):
        """Shape_Discriminator weights (scripts/discriminator.py:57-74) + the weight of its loss
        term (optimize.py:253).  ``state_dict=None`` switches the term off."""
        self.config_generation += 1     # whether the term runs and its weight are baked into captured launches
        with self._on_device():
            if state_dict is None:
                check(self.L.jrr_shape_critic_load(self.h, None, 0.0, _stream()), "jrr_shape_critic_load")
                return
            flat = flatten_shape_critic_state_dict(state_dict).to(self.device)
            check(self.L.jrr_shape_critic_load(self.h, _ptr(flat), float(w_shape), _stream()), "jrr_shape_critic_load")
            torch.cuda.current_stream().synchronize()   # `flat` may be freed after return

    def shape_critic_forward(self, betas):
        betas = _f32c(betas, "betas")
        B = betas.shape[0]
        out = torch.empty(B, device=self.device)
        with self._on_device():
            check(self.L.jrr_shape_critic_forward(self.h, B, _ptr(betas), _ptr(out), _stream()), "jrr_shape_critic_forward")
        self._done()
        return out

    # ------------------------------------------------------------------ critic training (8f-1)
    def critic_grad_accumulate(self, x6, target, G, loss=None, logical_batch=None, shape=False):
        """G += d/dparams mean((D(x) - target)^2) over these frames (divisor: logical batch)."""
        x = _f32c(x6, "x6" if not shape else "betas")
        B = x.shape[0]
        ws, wsz = self.workspace(B)
        fn = self.L.jrr_shape_critic_grad_accumulate if shape else self.L.jrr_critic_grad_accumulate
        with self._on_device():
            check(fn(self.h, B, int(logical_batch or B), _ptr(x), float(target), _ptr(G), _ptr(loss), ws, wsz,
                     _stream()), "jrr_critic_grad_accumulate")
        self._done()

    def critic_apply(self, params, G, m, v, t, lr, shape=False):
        fn = self.L.jrr_shape_critic_apply if shape else self.L.jrr_critic_apply
        with self._on_device():
            check(fn(self.h, _ptr(params), _ptr(G), _ptr(m), _ptr(v), _ptr(t), float(lr), _stream()), "jrr_critic_apply")
        self._done()

    # ------------------------------------------------------------------ SMPL forward / backward
    def smpl_forward(self, betas, pose, kind, want_verts=True, want_joints=True):
        B = pose.shape[0]
        betas, pose = _f32c(betas, "betas"), _f32c(pose, "pose")
        verts = torch.empty(B, 6890, 3, device=self.device) if want_verts else None
        joints = torch.empty(B, 49, 3, device=self.device) if want_joints else None
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_smpl_forward(self.h, B, _ptr(betas), _ptr(pose), kind, _ptr(verts),
                                          _ptr(joints), ws, wsz, _stream()), "jrr_smpl_forward")
        self._done()
        return verts, joints

    def smpl_backward(self, betas, pose, kind, dverts, djoints):
        B = pose.shape[0]
        betas, pose = _f32c(betas, "betas"), _f32c(pose, "pose")
        dverts = _f32c(dverts, "dvertices") if dverts is not None else None
        djoints = _f32c(djoints, "djoints") if djoints is not None else None
        dbetas = torch.empty(B, 10, device=self.device)
        dpose = torch.empty_like(pose)
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_smpl_backward(self.h, B, _ptr(betas), _ptr(pose), kind, _ptr(dverts),
                                           _ptr(djoints), _ptr(dbetas), _ptr(dpose), ws, wsz, _stream()),
                  "jrr_smpl_backward")
        self._done()
        return dbetas, dpose

    def find_joints(self, betas, pose, kind):
        B = pose.shape[0]
        betas, pose = _f32c(betas, "betas"), _f32c(pose, "pose")
        out = torch.empty(B, 17, 3, device=self.device)
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_find_joints(self.h, B, _ptr(betas), _ptr(pose), kind, _ptr(out), ws, wsz,
                                         _stream()), "jrr_find_joints")
        self._done()
        return out

    def find_joints_backward(self, betas, pose, kind, djoints):
        B = pose.shape[0]
        betas, pose, dj = _f32c(betas, "betas"), _f32c(pose, "pose"), _f32c(djoints, "djoints17")
        dbetas = torch.empty(B, 10, device=self.device)
        dpose = torch.empty_like(pose)
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_find_joints_backward(self.h, B, _ptr(betas), _ptr(pose), kind, _ptr(dj), _ptr(dbetas),
                                                  _ptr(dpose), ws, wsz, _stream()), "jrr_find_joints_backward")
        self._done()
        return dbetas, dpose

    def critic_forward(self, rot6d):
        B = rot6d.shape[0]
        x = _f32c(rot6d, "rot6d")
        out = torch.empty(B, 25, device=self.device)
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_critic_forward(self.h, B, _ptr(x), _ptr(out), ws, wsz, _stream()),
                  "jrr_critic_forward")
        self._done()
        return out

    # ------------------------------------------------------------------ refinement
    def refine_step(self, x6, betas, gt_mm, adam_m, adam_v, step_count, lr, w_joint, w_pose,
                    logical_batch=None, loss_out=None):
        """One fused iteration; x6/betas/adam_m/adam_v/step_count are updated in place."""
        B = x6.shape[0]
        for t, n in ((x6, "x6"), (betas, "betas"), (gt_mm, "gt_mm"), (adam_m, "adam_m"), (adam_v, "adam_v")):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise JrrError(f"{n} must be a contiguous fp32 CUDA tensor")
        LB = B if logical_batch is None else int(logical_batch)
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_refine_step(self.h, B, LB, _ptr(x6), _ptr(betas), _ptr(gt_mm), _ptr(adam_m),
                                         _ptr(adam_v), _ptr(step_count), lr, w_joint, w_pose,
                                         _ptr(loss_out), ws, wsz, _stream()), "jrr_refine_step")
        self._done()

    def refine_step_profiled(self, x6, betas, gt_mm, adam_m, adam_v, step_count, lr, w_joint, w_pose,
                             logical_batch=None, loss_out=None):
        """refine_step with per-kernel-group CUDA-event timing -> {name: ms} (synchronises)."""
        B = x6.shape[0]
        LB = B if logical_batch is None else int(logical_batch)
        ws, wsz = self.workspace(B)
        ms = (C.c_float * _lib.STEP_KERNELS)()
        with self._on_device():
            check(self.L.jrr_refine_step_profiled(self.h, B, LB, _ptr(x6), _ptr(betas), _ptr(gt_mm),
                                                  _ptr(adam_m), _ptr(adam_v), _ptr(step_count), lr, w_joint,
                                                  w_pose, _ptr(loss_out), ws, wsz, _stream(), ms),
                  "jrr_refine_step_profiled")
        self._done()
        return {self.L.jrr_step_kernel_name(i).decode(): float(ms[i]) for i in range(_lib.STEP_KERNELS)}

    def camera_fit(self, x6, betas, gt_j2d, cam, iters=1000, lr=1e-2, logical_batch=None, loss_out=None):
        """optimize.py:187-199: Adam on the camera translation only; cam [B,3] is updated in place."""
        B = x6.shape[0]
        LB = B if logical_batch is None else int(logical_batch)
        x6, betas, gt_j2d = _f32c(x6, "x6"), _f32c(betas, "betas"), _f32c(gt_j2d, "gt_j2d")
        if not (cam.is_cuda and cam.dtype == torch.float32 and cam.is_contiguous()):
            raise JrrError("cam must be a contiguous fp32 CUDA tensor")
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_camera_fit(self.h, B, LB, _ptr(x6), _ptr(betas), _ptr(gt_j2d), _ptr(cam), int(iters),
                                        lr, _ptr(loss_out), ws, wsz, _stream()), "jrr_camera_fit")
        self._done()

    def refine_step_2d(self, x6, betas, gt_mm, gt_j2d, cam, adam_m, adam_v, cam_m, cam_v, step_count, lr, w_joint,
                       w_pose, w_2d, logical_batch=None, loss_out=None):
        B = x6.shape[0]
        LB = B if logical_batch is None else int(logical_batch)
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_refine_step_2d(self.h, B, LB, _ptr(x6), _ptr(betas), _ptr(gt_mm), _ptr(gt_j2d), _ptr(cam),
                                            _ptr(adam_m), _ptr(adam_v), _ptr(cam_m), _ptr(cam_v), _ptr(step_count),
                                            lr, w_joint, w_pose, w_2d, _ptr(loss_out), ws, wsz, _stream()),
                  "jrr_refine_step_2d")
        self._done()

    def set_external_gradient(self, dx6=None, dbetas=None, dcam=None):
        """Gradients of loss terms computed outside the refinement step (the silhouette term), added before Adam by
        every later refine_step / refine_step_2d until cleared (call without arguments)."""
        self._ext = (dx6, dbetas, dcam)        # keep the tensors alive while the library holds their pointers
        check(self.L.jrr_set_external_gradient(self.h, _ptr(dx6), _ptr(dbetas), _ptr(dcam)), "jrr_set_external_gradient")

    def regressor_grad_accumulate(self, x6, betas, gt_mm, G_accum, loss_accum, logical_batch=None):
        B = x6.shape[0]
        LB = B if logical_batch is None else int(logical_batch)
        x6, betas, gt_mm = _f32c(x6, "x6"), _f32c(betas, "betas"), _f32c(gt_mm, "gt_mm")
        ws, wsz = self.workspace(B)
        with self._on_device():
            check(self.L.jrr_regressor_grad_accumulate(self.h, B, LB, _ptr(x6), _ptr(betas), _ptr(gt_mm),
                                                       _ptr(G_accum), _ptr(loss_accum), ws, wsz, _stream()),
                  "jrr_regressor_grad_accumulate")
        self._done()

    def regressor_apply(self, J17_raw, mask, G_accum, adam_m, adam_v, step_count, lr):
        with self._on_device():
            check(self.L.jrr_regressor_apply(self.h, _ptr(J17_raw), _ptr(mask), _ptr(G_accum), _ptr(adam_m),
                                             _ptr(adam_v), _ptr(step_count), lr, _stream()),
                  "jrr_regressor_apply")
        self._done()

    # ------------------------------------------------------------------ diagnostics
    def debug_gemm(self, A, B, impl=0):
        """C[M,N] = A[M,K] @ B[N,K]^T through the 3xTF32 kernel (impl 0) or the SIMT one (1)."""
        M, K = A.shape
        N = B.shape[0]
        A, B = _f32c(A, "A"), _f32c(B, "B")
        Cc = torch.empty(M, N, device=self.device)
        scratch = torch.empty(2 * (M + N) * K, device=self.device)
        with self._on_device():
            check(self.L.jrr_debug_gemm(self.h, impl, M, N, K, _ptr(A), _ptr(B), _ptr(Cc), _ptr(scratch),
                                        _stream()), "jrr_debug_gemm")
        self._done()
        return Cc
