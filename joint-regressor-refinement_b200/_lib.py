"""ctypes binding of libjrr.so (include/jrr.h).  There is no fallback: if the CUDA library
is missing or does not load, importing a native entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjrr.so")

POSE_ROTMAT, POSE_AXIS_ANGLE, POSE_ROT6D = 0, 1, 2
CRITIC_PARAMS = 1840153
STEP_KERNELS = 14

EXPORTS = [
    "jrr_last_error", "jrr_abi_version", "jrr_model_create", "jrr_model_destroy",
    "jrr_set_regressor", "jrr_critic_load", "jrr_workspace_bytes", "jrr_smpl_forward",
    "jrr_smpl_backward", "jrr_find_joints", "jrr_critic_forward", "jrr_refine_step",
    "jrr_regressor_grad_accumulate", "jrr_regressor_apply", "jrr_last_launch_count",
    "jrr_debug_gemm", "jrr_refine_step_profiled", "jrr_step_kernel_name", "jrr_camera_fit", "jrr_refine_step_2d", "jrr_evaluate",
    "jrr_shape_critic_load", "jrr_shape_critic_forward", "jrr_critic_grad_accumulate", "jrr_critic_apply",
    "jrr_shape_critic_grad_accumulate", "jrr_shape_critic_apply", "jrr_set_loss_path", "jrr_debug_tma_probe", "jrr_debug_set_gemm_prof",
    "jrr_critic_layer2_bwd_products", "jrr_silhouette_workspace_bytes", "jrr_silhouette_forward", "jrr_silhouette_backward", "jrr_set_external_gradient", "jrr_find_joints_backward",
]


class JrrModelDesc(C.Structure):
    _fields_ = [
        ("v_template_host", C.c_void_p), ("shapedirs_host", C.c_void_p),
        ("posedirs_host", C.c_void_p), ("J_regressor_host", C.c_void_p),
        ("parents_host", C.c_void_p), ("lbs_weights_host", C.c_void_p),
        ("J_regressor_extra_host", C.c_void_p), ("joint_map_host", C.c_void_p),
        ("vertex_picks_host", C.c_void_p), ("device", C.c_int32), ("gemm_impl", C.c_int32),
    ]


class JrrError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JrrError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                       "g.build()'` -- there is no CPU or PyTorch fallback for the hot path")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
    L.jrr_last_error.restype = C.c_char_p
    L.jrr_abi_version.restype = C.c_int
    L.jrr_model_create.argtypes = [C.POINTER(JrrModelDesc), C.POINTER(vp)]
    L.jrr_model_destroy.argtypes = [vp]
    L.jrr_set_regressor.argtypes = [vp, vp, vp, vp]
    L.jrr_critic_load.argtypes = [vp, vp, vp]
    L.jrr_set_loss_path.argtypes = [vp, C.c_int, vp]
    L.jrr_critic_layer2_bwd_products.argtypes = [vp, C.c_int64]
    L.jrr_shape_critic_load.argtypes = [vp, vp, C.c_float, vp]
    L.jrr_shape_critic_forward.argtypes = [vp, i64, vp, vp, vp]
    L.jrr_critic_grad_accumulate.argtypes = [vp, i64, i64, vp, C.c_float, vp, vp, vp, sz, vp]
    L.jrr_critic_apply.argtypes = [vp, vp, vp, vp, vp, vp, C.c_float, vp]
    L.jrr_shape_critic_grad_accumulate.argtypes = [vp, i64, i64, vp, C.c_float, vp, vp, vp, sz, vp]
    L.jrr_shape_critic_apply.argtypes = [vp, vp, vp, vp, vp, vp, C.c_float, vp]
    L.jrr_workspace_bytes.argtypes = [vp, i64]
    L.jrr_workspace_bytes.restype = sz
    L.jrr_smpl_forward.argtypes = [vp, i64, vp, vp, C.c_int, vp, vp, vp, sz, vp]
    L.jrr_smpl_backward.argtypes = [vp, i64, vp, vp, C.c_int, vp, vp, vp, vp, vp, sz, vp]
    L.jrr_find_joints.argtypes = [vp, i64, vp, vp, C.c_int, vp, vp, sz, vp]
    L.jrr_find_joints_backward.argtypes = [vp, i64, vp, vp, C.c_int, vp, vp, vp, vp, sz, vp]
    L.jrr_critic_forward.argtypes = [vp, i64, vp, vp, vp, sz, vp]
    L.jrr_refine_step.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, sz, vp]
    L.jrr_regressor_grad_accumulate.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, sz, vp]
    L.jrr_regressor_apply.argtypes = [vp, vp, vp, vp, vp, vp, vp, f32, vp]
    L.jrr_last_launch_count.restype = i64
    L.jrr_refine_step_profiled.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, sz, vp, vp]
    L.jrr_step_kernel_name.argtypes = [C.c_int]
    L.jrr_step_kernel_name.restype = C.c_char_p
    L.jrr_camera_fit.argtypes = [vp, i64, i64, vp, vp, vp, vp, C.c_int, f32, vp, vp, sz, vp]
    L.jrr_refine_step_2d.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, f32, f32, f32, vp, vp, sz, vp]
    L.jrr_evaluate.argtypes = [i64, vp, vp, vp, vp, vp, sz, vp]
    L.jrr_set_external_gradient.argtypes = [vp, vp, vp, vp]
    L.jrr_silhouette_workspace_bytes.argtypes = [i64, i64, i64, C.c_int]
    L.jrr_silhouette_workspace_bytes.restype = sz
    L.jrr_silhouette_forward.argtypes = [i64, vp, i64, vp, vp, i64, C.c_int, f32, f32, C.c_int, vp, i64, vp, vp, vp, vp, sz, vp]
    L.jrr_silhouette_backward.argtypes = [i64, vp, i64, vp, vp, i64, vp, vp, C.c_int, f32, f32, C.c_int, vp, vp, vp, vp, i64, f32,
                                          vp, vp, vp, sz, vp]
    L.jrr_debug_gemm.argtypes = [vp, C.c_int, i64, i64, i64, vp, vp, vp, vp, vp]
    L.jrr_debug_set_gemm_prof.argtypes = [vp, C.c_int]
    L.jrr_debug_tma_probe.argtypes = [vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("jrr_abi_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().jrr_last_error().decode("utf-8", "replace")
        raise JrrError(f"{what} failed (status {rc}): {msg}")
