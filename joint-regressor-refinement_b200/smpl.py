"""Drop-in for the reference body model, ``scripts/smpl.py:61-85`` (a subclass of
``smplx.SMPL``): same constructor, buffers, forward signature and output object, with the
arithmetic done by hand-written sm_100a kernels behind a ``torch.autograd.Function``.

    smpl = SMPL(model_dict=make_smpl_model()).cuda()           # or SMPL("path/to/model_dir")
    out = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    out.vertices  # [B,6890,3]     out.joints  # [B,49,3]
"""
from __future__ import annotations

import os
import pickle
from collections import namedtuple

import numpy as np
import torch
from torch import nn

from . import synthetic
from ._lib import POSE_AXIS_ANGLE, POSE_ROTMAT, JrrError
from .native import NativeModel

SMPLOutput = namedtuple("SMPLOutput", ["vertices", "joints", "full_pose", "betas", "global_orient",
                                       "body_pose"])
SMPLOutput.__new__.__defaults__ = (None,) * 6

VIBE_DATA_DIR = "data/vibe_data"
JOINT_REGRESSOR_TRAIN_EXTRA = os.path.join(VIBE_DATA_DIR, "J_regressor_extra.npy")


class SMPLFunction(torch.autograd.Function):
    """(betas [B,10], pose [B,24,*]) -> (vertices [B,6890,3], joints [B,49,3]).
    Backward = jrr_smpl_backward (recomputes the forward intermediates)."""

    @staticmethod
    def forward(ctx, betas, pose, native, kind):
        verts, joints = native.smpl_forward(betas, pose, kind, True, True)
        ctx.save_for_backward(betas, pose)
        ctx.native, ctx.kind = native, kind
        return verts, joints

    @staticmethod
    def backward(ctx, dverts, djoints):
        betas, pose = ctx.saved_tensors
        if dverts is None and djoints is None:
            return torch.zeros_like(betas), torch.zeros_like(pose), None, None
        dbetas, dpose = ctx.native.smpl_backward(betas, pose, ctx.kind, dverts, djoints)
        return dbetas.view_as(betas), dpose.view_as(pose), None, None


def _load_model_file(model_path: str, gender: str) -> dict:
    """smplx-style lookup: a file, or a directory holding SMPL_{GENDER}.pkl / .npz."""
    path = model_path
    if os.path.isdir(path):
        for ext in ("pkl", "npz"):
            cand = os.path.join(path, f"SMPL_{gender.upper()}.{ext}")
            if os.path.exists(cand):
                path = cand
                break
        else:
            raise FileNotFoundError(f"no SMPL_{gender.upper()}.pkl/.npz under {model_path}")
    if path.endswith(".npz"):
        data = dict(np.load(path, allow_pickle=True))
    else:
        with open(path, "rb") as f:
            data = pickle.load(f, encoding="latin1")   # official files need chumpy installed
    dense = lambda a: np.asarray(a.todense() if hasattr(a, "todense") else a)
    parents = np.asarray(data["kintree_table"])[0].astype(np.int64) if "kintree_table" in data \
        else np.asarray(data["parents"]).astype(np.int64)
    parents[0] = -1
    out = {
        "v_template": np.asarray(data["v_template"], dtype=np.float32),
        "shapedirs": np.asarray(data["shapedirs"], dtype=np.float32)[:, :, :10],
        "J_regressor": dense(data["J_regressor"]).astype(np.float32),
        "lbs_weights": np.asarray(data.get("weights", data.get("lbs_weights")), dtype=np.float32),
        "parents": parents,
        "faces": np.asarray(data.get("f", data.get("faces"))).astype(np.int64),
    }
    pd = np.asarray(data["posedirs"], dtype=np.float32)
    out["posedirs"] = pd.reshape(-1, pd.shape[-1]).T.copy() if pd.ndim == 3 else pd
    return out


class SMPL(nn.Module):
    """Extension of the SMPL body model to 49 joints (scripts/smpl.py:61-85)."""

    NUM_JOINTS = 23
    NUM_BODY_JOINTS = 23

    def __init__(self, model_path=None, batch_size=1, create_transl=True, gender="neutral",
                 model_dict=None, J_regressor_extra=None, gemm_impl=0, **kwargs):
        super().__init__()
        if model_dict is None:
            if model_path is None:
                raise ValueError("give model_path (SMPL .pkl/.npz) or model_dict")
            model_dict = _load_model_file(model_path, gender)
        md = dict(model_dict)
        if J_regressor_extra is not None:
            md["J_regressor_extra"] = np.asarray(J_regressor_extra, dtype=np.float32)
        elif "J_regressor_extra" not in md:
            md["J_regressor_extra"] = np.load(JOINT_REGRESSOR_TRAIN_EXTRA).astype(np.float32)
        md.setdefault("joint_map", synthetic.JOINT_MAP_49)
        md.setdefault("vertex_picks", synthetic.VERTEX_PICKS)
        self._model_np = md
        self._gemm_impl = gemm_impl
        self._native = None
        self.batch_size = batch_size
        t = lambda k: torch.as_tensor(np.asarray(md[k], dtype=np.float32))
        self.register_buffer("v_template", t("v_template"))
        self.register_buffer("shapedirs", t("shapedirs"))
        self.register_buffer("posedirs", t("posedirs"))
        self.register_buffer("J_regressor", t("J_regressor"))
        self.register_buffer("lbs_weights", t("lbs_weights"))
        self.register_buffer("J_regressor_extra", t("J_regressor_extra"))
        self.register_buffer("parents", torch.as_tensor(np.asarray(md["parents"], dtype=np.int64)))
        self.register_buffer("faces_tensor", torch.as_tensor(np.asarray(md["faces"], dtype=np.int64)))
        self.faces = np.asarray(md["faces"])
        self.joint_map = torch.as_tensor(np.asarray(md["joint_map"], dtype=np.int64))
        # smpl.py:75-76 regresses the 9 extra joints from the TRANSLATED vertices, so `transl` reaches them scaled by
        # their regressor row sums (1 only when the rows are normalised); the other 45 joints get it unscaled
        jm = np.asarray(md["joint_map"], dtype=np.int64)
        rs = np.asarray(md["J_regressor_extra"], dtype=np.float64).sum(axis=1)
        self.register_buffer("_transl_scale", torch.as_tensor(
            np.where(jm >= 45, rs[np.clip(jm - 45, 0, len(rs) - 1)], 1.0).astype(np.float32)), persistent=False)
        # smplx default parameters (batch_size rows)
        self.betas = nn.Parameter(torch.zeros(batch_size, 10))
        self.global_orient = nn.Parameter(torch.zeros(batch_size, 3))
        self.body_pose = nn.Parameter(torch.zeros(batch_size, 69))
        if create_transl:
            self.transl = nn.Parameter(torch.zeros(batch_size, 3))
        else:
            self.register_parameter("transl", None)

    # the packed device copy is created lazily on the device the buffers live on
    def native(self) -> NativeModel:
        dev = self.v_template.device
        if dev.type != "cuda":
            raise JrrError("SMPL must be moved to a CUDA device before use (no CPU path)")
        if self._native is None or self._native.device != torch.device("cuda", dev.index or 0):
            self._native = NativeModel(self._model_np, dev, self._gemm_impl)
        return self._native

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None,
                return_verts=True, return_full_pose=False, pose2rot=True, **kwargs):
        betas = self.betas if betas is None else betas
        body_pose = self.body_pose if body_pose is None else body_pose
        global_orient = self.global_orient if global_orient is None else global_orient
        if transl is None and self.transl is not None:
            transl = self.transl
        B = max(betas.shape[0], body_pose.shape[0], global_orient.shape[0])
        if betas.shape[0] != B:
            betas = betas.expand(B, -1)
        if pose2rot:
            full = torch.cat([global_orient.reshape(-1, 1, 3).expand(B, -1, -1),
                              body_pose.reshape(-1, 23, 3).expand(B, -1, -1)], dim=1)
            kind = POSE_AXIS_ANGLE
        else:
            full = torch.cat([global_orient.reshape(-1, 1, 3, 3).expand(B, -1, -1, -1),
                              body_pose.reshape(-1, 23, 3, 3).expand(B, -1, -1, -1)], dim=1).reshape(B, 24, 9)
            kind = POSE_ROTMAT
        verts, joints = SMPLFunction.apply(betas.float().contiguous(), full.float().contiguous(),
                                           self.native(), kind)
        if transl is not None:
            if transl.shape[0] != B:
                transl = transl.expand(B, -1)
            verts = verts + transl[:, None]
            joints = joints + transl[:, None] * self._transl_scale[None, :, None]
        return SMPLOutput(vertices=verts if return_verts else None, joints=joints, betas=betas,
                          global_orient=global_orient, body_pose=body_pose,
                          full_pose=full if return_full_pose else None)


def get_smpl_faces(model_dir=VIBE_DATA_DIR):
    """scripts/smpl.py:88-90."""
    return SMPL(model_dir, batch_size=1, create_transl=False).faces
