"""Seeded synthetic SMPL-shaped body model and refinement inputs.

SMPL model files, SPIN estimates and Human3.6M are not available offline, so
the hot path is exercised on a random-init model of the canonical shape
(6890 vertices, 24 joints, 10 betas, 207 pose-blend dims).
Everything here is numpy (PCG64 is platform-stable), so the GPU box regenerates
bit-identical constants from the seed and nothing large has to be committed.

The buffers mirror what ``smplx.SMPL`` registers and ``scripts/smpl.py:61-70``
adds (``J_regressor_extra``, ``joint_map``).
"""
from __future__ import annotations

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_FEATS = 207
NUM_EXTRA = 9
NUM_H36M = 17

# canonical SMPL kinematic tree
PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14,
                    16, 17, 18, 19, 20, 21], dtype=np.int64)

# approximate rest-pose joint centres of a 1.7 m body, metres, y up
_REST = np.array([
    [0.00, -0.22, 0.03], [0.07, -0.31, 0.02], [-0.07, -0.31, 0.02],
    [0.00, -0.11, 0.00], [0.10, -0.69, 0.02], [-0.10, -0.69, 0.02],
    [0.00, 0.03, 0.02], [0.09, -1.09, -0.02], [-0.09, -1.09, -0.02],
    [0.00, 0.08, 0.02], [0.11, -1.15, 0.10], [-0.11, -1.15, 0.10],
    [0.00, 0.29, -0.01], [0.08, 0.20, 0.00], [-0.08, 0.20, 0.00],
    [0.00, 0.37, 0.03], [0.17, 0.23, -0.01], [-0.17, 0.23, -0.01],
    [0.43, 0.22, -0.03], [-0.43, 0.22, -0.03], [0.68, 0.22, -0.02],
    [-0.68, 0.22, -0.02], [0.77, 0.21, -0.03], [-0.77, 0.21, -0.03],
], dtype=np.float64)

# limb radius used to scatter surface points around each bone (child joint index)
_RADIUS = np.array([0.12, 0.09, 0.09, 0.13, 0.06, 0.06, 0.13, 0.04, 0.04, 0.13,
                    0.035, 0.035, 0.06, 0.07, 0.07, 0.09, 0.05, 0.05, 0.04,
                    0.04, 0.03, 0.03, 0.035, 0.035], dtype=np.float64)

# smplx VertexJointSelector ids: nose, eyes, ears, feet, finger tips
VERTEX_PICKS = np.array([332, 6260, 2800, 4071, 583,
                         3216, 3226, 3387, 6617, 6624, 6787,
                         2746, 2319, 2445, 2556, 2673,
                         6191, 5782, 5905, 6016, 6133], dtype=np.int64)

# scripts/smpl.py:12-49,66 -- the 49 names gathered from the 54-joint stack
_JOINT_MAP_54 = {
    'OP Nose': 24, 'OP Neck': 12, 'OP RShoulder': 17, 'OP RElbow': 19, 'OP RWrist': 21,
    'OP LShoulder': 16, 'OP LElbow': 18, 'OP LWrist': 20, 'OP MidHip': 0, 'OP RHip': 2,
    'OP RKnee': 5, 'OP RAnkle': 8, 'OP LHip': 1, 'OP LKnee': 4, 'OP LAnkle': 7,
    'OP REye': 25, 'OP LEye': 26, 'OP REar': 27, 'OP LEar': 28, 'OP LBigToe': 29,
    'OP LSmallToe': 30, 'OP LHeel': 31, 'OP RBigToe': 32, 'OP RSmallToe': 33, 'OP RHeel': 34,
    'Right Ankle': 8, 'Right Knee': 5, 'Right Hip': 45, 'Left Hip': 46, 'Left Knee': 4,
    'Left Ankle': 7, 'Right Wrist': 21, 'Right Elbow': 19, 'Right Shoulder': 17,
    'Left Shoulder': 16, 'Left Elbow': 18, 'Left Wrist': 20, 'Neck (LSP)': 47,
    'Top of Head (LSP)': 48, 'Pelvis (MPII)': 49, 'Thorax (MPII)': 50, 'Spine (H36M)': 51,
    'Jaw (H36M)': 52, 'Head (H36M)': 53, 'Nose': 24, 'Left Eye': 26, 'Right Eye': 25,
    'Left Ear': 28, 'Right Ear': 27,
}
_JOINT_NAMES_49 = [
    'OP Nose', 'OP Neck', 'OP RShoulder', 'OP RElbow', 'OP RWrist', 'OP LShoulder',
    'OP LElbow', 'OP LWrist', 'OP MidHip', 'OP RHip', 'OP RKnee', 'OP RAnkle', 'OP LHip',
    'OP LKnee', 'OP LAnkle', 'OP REye', 'OP LEye', 'OP REar', 'OP LEar', 'OP LBigToe',
    'OP LSmallToe', 'OP LHeel', 'OP RBigToe', 'OP RSmallToe', 'OP RHeel', 'Right Ankle',
    'Right Knee', 'Right Hip', 'Left Hip', 'Left Knee', 'Left Ankle', 'Right Wrist',
    'Right Elbow', 'Right Shoulder', 'Left Shoulder', 'Left Elbow', 'Left Wrist',
    'Neck (LSP)', 'Top of Head (LSP)', 'Pelvis (MPII)', 'Thorax (MPII)', 'Spine (H36M)',
    'Jaw (H36M)', 'Head (H36M)', 'Nose', 'Left Eye', 'Right Eye', 'Left Ear', 'Right Ear',
]
JOINT_MAP_49 = np.array([_JOINT_MAP_54[n] for n in _JOINT_NAMES_49], dtype=np.int64)


def _sparse_rows(rng, targets, verts, nnz):
    """Rows of positive weights on the `nnz` vertices nearest each target, rows sum to 1."""
    out = np.zeros((targets.shape[0], verts.shape[0]), dtype=np.float64)
    for r, t in enumerate(targets):
        d = np.linalg.norm(verts - t[None], axis=1)
        idx = np.argsort(d, kind="stable")[:nnz]
        w = rng.uniform(0.2, 1.0, size=nnz)
        out[r, idx] = w / w.sum()
    return out


def make_smpl_model(seed: int = 0, skin_weights_per_vertex: int = 4) -> dict:
    """Random-init SMPL of the canonical shape.  Returns float32/int64 numpy arrays with
    the smplx buffer names: v_template, shapedirs, posedirs, J_regressor, parents,
    lbs_weights, faces, J_regressor_extra, joint_map, vertex_picks.
    ``skin_weights_per_vertex`` (SMPL: 4) sets the non-zeros of every ``lbs_weights`` row; larger values
    exercise the multi-pass skinning of models that are not 4-sparse."""
    rng = np.random.default_rng(seed)
    # vertices are generated bone by bone so that neighbouring ids are spatial
    # neighbours, as in the real SMPL mesh
    bones = [(int(PARENTS[j]), j) for j in range(1, NUM_JOINTS)]
    lens = np.array([np.linalg.norm(_REST[c] - _REST[p]) + 0.05 for p, c in bones])
    area = lens * np.array([_RADIUS[c] for _, c in bones])
    counts = np.floor(area / area.sum() * (NUM_VERTS - 400)).astype(int)
    head_extra = NUM_VERTS - counts.sum()
    verts = []
    for (p, c), n in zip(bones, counts):
        t = np.sort(rng.uniform(-0.1, 1.1, size=n))
        centre = _REST[p][None] * (1 - t[:, None]) + _REST[c][None] * t[:, None]
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        verts.append(centre + d * _RADIUS[c] * rng.uniform(0.8, 1.0, size=(n, 1)))
    # a head blob above joint 15
    d = rng.normal(size=(head_extra, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    verts.append(_REST[15][None] + np.array([0, 0.09, 0.0])[None] + d * 0.1)
    v_template = np.concatenate(verts, axis=0)
    assert v_template.shape == (NUM_VERTS, 3)

    shapedirs = rng.normal(scale=0.01, size=(NUM_VERTS, 3, NUM_BETAS))
    posedirs = rng.normal(scale=0.002, size=(NUM_POSE_FEATS, NUM_VERTS * 3))

    J_regressor = _sparse_rows(rng, _REST, v_template, 32)

    # exactly 4 (default) non-zeros per vertex: softmax(-distance) over the nearest joints
    dist = np.linalg.norm(v_template[:, None, :] - _REST[None, :, :], axis=2)
    near = np.argsort(dist, axis=1, kind="stable")[:, :skin_weights_per_vertex]
    logits = -np.take_along_axis(dist, near, axis=1) / 0.05
    logits -= logits.max(axis=1, keepdims=True)
    w4 = np.exp(logits)
    w4 /= w4.sum(axis=1, keepdims=True)
    w4 = np.maximum(w4, 1e-4)
    w4 /= w4.sum(axis=1, keepdims=True)
    lbs_weights = np.zeros((NUM_VERTS, NUM_JOINTS))
    np.put_along_axis(lbs_weights, near, w4, axis=1)

    extra_targets = np.stack([
        _REST[2] + [0.0, 0.03, 0.0], _REST[1] + [0.0, 0.03, 0.0],      # R/L hip (LSP)
        _REST[12] + [0.0, 0.02, 0.0], _REST[15] + [0.0, 0.17, 0.0],     # neck, head top
        _REST[0] + [0.0, 0.02, 0.0], _REST[9] + [0.0, 0.10, 0.0],       # pelvis, thorax
        _REST[6], _REST[15] + [0.0, -0.02, 0.06], _REST[15] + [0.0, 0.06, 0.0],
    ])
    J_regressor_extra = _sparse_rows(rng, extra_targets, v_template, 16)

    faces = rng.integers(0, NUM_VERTS, size=(13776, 3), dtype=np.int64)

    return {
        "v_template": v_template.astype(np.float32),
        "shapedirs": shapedirs.astype(np.float32),
        "posedirs": posedirs.astype(np.float32),
        "J_regressor": J_regressor.astype(np.float32),
        "parents": PARENTS.copy(),
        "lbs_weights": lbs_weights.astype(np.float32),
        "faces": faces,
        "J_regressor_extra": J_regressor_extra.astype(np.float32),
        "joint_map": JOINT_MAP_49.copy(),
        "vertex_picks": VERTEX_PICKS.copy(),
    }


def make_local_faces(v_template: np.ndarray, n_faces: int = 13776, lbs_weights: np.ndarray | None = None) -> np.ndarray:
    """A surface-like triangle soup for the silhouette term: the model's own ``faces`` are random vertex triples (nothing on
    the hot path reads them), which would be body-sized triangles.  Here every vertex spans triangles with its nearest
    neighbours (2 per vertex: neighbours 1-2 and 3-4), so faces have edges of 1-2 cm like SMPL's 13 776 (two to three pixels at
    224 x 224; the rasteriser's cost goes with the pixel area of the faces' bounding boxes).  With
    ``lbs_weights`` the neighbours are taken among the vertices of the same body part (dominant skinning joint), so no face
    bridges two parts that merely touch in the rest pose (hand / thigh, arm / torso) and stretches across the image once
    the body is posed -- a real mesh has no such faces."""
    from scipy.spatial import cKDTree
    v = np.asarray(v_template, dtype=np.float64)
    n = v.shape[0]
    part = np.zeros(n, dtype=np.int64) if lbs_weights is None else np.asarray(lbs_weights).argmax(axis=1)
    nb = np.zeros((n, 9), dtype=np.int64)
    _, nb_all = cKDTree(v).query(v, k=9)
    for p in np.unique(part):
        idx = np.nonzero(part == p)[0]
        if idx.size < 9:
            nb[idx] = nb_all[idx]           # (a part too small for its own neighbourhood)
            continue
        _, loc = cKDTree(v[idx]).query(v[idx], k=9)
        nb[idx] = idx[loc]
    a = np.stack([nb[:, 0], nb[:, 1], nb[:, 2]], axis=1)
    b = np.stack([nb[:, 0], nb[:, 3], nb[:, 4]], axis=1)
    faces = np.concatenate([a, b], axis=0)
    return np.ascontiguousarray(faces[:n_faces], dtype=np.int64)


def make_dense_regressor(seed: int = 0) -> np.ndarray:
    """Dense 17x6890 raw regressor |N(0,1)| (exercises the dense reduction)."""
    rng = np.random.default_rng(seed + 7919)
    return np.abs(rng.normal(size=(NUM_H36M, NUM_VERTS))).astype(np.float32)


def make_sparse_regressor(model: dict, seed: int = 0) -> np.ndarray:
    """Stand-in with the sparsity profile of models/retrained_J_Regressor.pt (about
    6 positive + a few negative entries per row, raw row sums not normalised) for
    boxes where the reference artefact is absent."""
    rng = np.random.default_rng(seed + 104729)
    targets = _REST[[0, 1, 4, 7, 2, 5, 8, 6, 12, 15, 15, 16, 18, 20, 17, 19, 21]].copy()
    targets[9] += [0.0, 0.02, 0.08]
    targets[10] += [0.0, 0.12, 0.0]
    pos = _sparse_rows(rng, targets, model["v_template"].astype(np.float64), 4)
    pos *= rng.uniform(0.8, 2.1, size=(NUM_H36M, 1))
    neg_idx = rng.integers(0, NUM_VERTS, size=(NUM_H36M, 3))
    for r in range(NUM_H36M):
        for c in neg_idx[r]:
            if pos[r, c] == 0:
                pos[r, c] = -rng.uniform(0.01, 0.2)
    return pos.astype(np.float32)


def axis_angle_to_rotmat(aa: np.ndarray) -> np.ndarray:
    """Plain Rodrigues (float64), used only to synthesise inputs."""
    aa = np.asarray(aa, dtype=np.float64)
    ang = np.linalg.norm(aa, axis=-1, keepdims=True)
    n = aa / np.maximum(ang, 1e-12)
    x, y, z = n[..., 0], n[..., 1], n[..., 2]
    zero = np.zeros_like(x)
    K = np.stack([zero, -z, y, z, zero, -x, -y, x, zero], axis=-1).reshape(aa.shape[:-1] + (3, 3))
    s = np.sin(ang)[..., None]
    c = np.cos(ang)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def rotmat_to_rot6d(R: np.ndarray) -> np.ndarray:
    """First two columns in the interleaved ``view(-1,3,2)`` layout of
    scripts/utils.py:198-200: (R00,R01,R10,R11,R20,R21)."""
    return R[..., :, :2].reshape(R.shape[:-2] + (6,))


def make_pose_inputs(n_frames: int, seed: int = 0) -> dict:
    """Model-independent part of the synthetic frames: 'true' pose/shape and the
    perturbed initial estimate ("SPIN").  float32 numpy arrays:
    true_rotmat [N,24,3,3], true_betas [N,10], x6 [N,24,6], betas [N,10], gt_noise [N,17,3] (mm)."""
    rng = np.random.default_rng(seed)
    aa = rng.normal(scale=0.3, size=(n_frames, NUM_JOINTS, 3))
    aa[:, 0] = rng.normal(scale=1.0, size=(n_frames, 3))
    true_betas = np.clip(rng.normal(size=(n_frames, NUM_BETAS)), -3, 3)
    aa0 = aa + rng.normal(scale=0.15, size=aa.shape)
    x6 = rotmat_to_rot6d(axis_angle_to_rotmat(aa0)) + rng.normal(scale=0.01, size=(n_frames, NUM_JOINTS, 6))
    betas0 = true_betas + rng.normal(scale=0.5, size=true_betas.shape)
    gt_noise = rng.normal(scale=5.0, size=(n_frames, NUM_H36M, 3))
    return {
        "true_rotmat": axis_angle_to_rotmat(aa).astype(np.float32),
        "true_betas": true_betas.astype(np.float32),
        "x6": x6.astype(np.float32),
        "betas": betas0.astype(np.float32),
        "gt_noise": gt_noise.astype(np.float32),
    }
