"""A SECOND, independent CPU statement of the SMPL forward -- TEST INFRASTRUCTURE ONLY.

Why it exists: the arithmetic of ``scripts/smpl.py:72-78`` lives in ``smplx==0.1.26``
(``requirements.txt:12``), which is neither vendored nor installable offline, so
``oracle/jrr_oracle.py::lbs`` is a restatement that cannot be pinned against the real package
(PARITY UNPINNED).  This file removes the "one author, one formulation" risk as far as that is
possible offline: it is written from the SMPL paper (Loper et al. 2015, eqs. 2-10) in a different
form and shares no code with ``jrr_oracle.lbs``:

* plain NumPy fp64, one pose at a time, an explicit Python loop over vertices and joints;
* rotations by the matrix exponential of the skew matrix (``scipy.linalg.expm``), not the closed
  Rodrigues formula;
* 4x4 homogeneous world transforms built by recursion from the root, and the skinning transform as
  ``G_k(theta, J) @ inv(G_k(0, J))`` with a numerically inverted rest-pose matrix (eq. 4), not the
  "subtract the rotated rest joint" shortcut;
* blend shapes applied per vertex with explicit sums over the 10 shape and 207 pose coefficients.

``tests/test_oracle.py::test_independent_lbs_statement_agrees`` checks the two statements against
each other (vertices and joints, 1e-9 relative in fp64).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm


def _skew(r):
    return np.array([[0.0, -r[2], r[1]], [r[2], 0.0, -r[0]], [-r[1], r[0], 0.0]])


def rotation_from_axis_angle(r):
    """exp([r]x): the rotation by |r| radians about r/|r| (SMPL paper eq. 1 states its closed form)."""
    return expm(_skew(np.asarray(r, dtype=np.float64)))


def _world_transform(k, parents, local, memo):
    """G_k = G_parent(k) @ local_k, by recursion from the root (SMPL paper eq. 3)."""
    if k in memo:
        return memo[k]
    p = int(parents[k])
    G = local[k] if p < 0 else _world_transform(p, parents, local, memo) @ local[k]
    memo[k] = G
    return G


def _chain(rot, joints, parents):
    """World transforms of the 24 joints for rotations `rot` [24,3,3] and rest joint positions [24,3]."""
    local = []
    for k in range(24):
        p = int(parents[k])
        M = np.eye(4)
        M[:3, :3] = rot[k]
        M[:3, 3] = joints[k] if p < 0 else joints[k] - joints[p]
        local.append(M)
    memo = {}
    return [_world_transform(k, parents, local, memo) for k in range(24)]


def smpl_forward_one(model: dict, betas, rot):
    """One pose.  betas [10], rot [24,3,3] -> (vertices [6890,3], posed joints [24,3])."""
    vt = np.asarray(model["v_template"], dtype=np.float64)
    S = np.asarray(model["shapedirs"], dtype=np.float64)            # [V,3,10]
    P = np.asarray(model["posedirs"], dtype=np.float64)             # [207, 3V]
    Jr = np.asarray(model["J_regressor"], dtype=np.float64)         # [24,V]
    W = np.asarray(model["lbs_weights"], dtype=np.float64)          # [V,24]
    parents = np.asarray(model["parents"])
    betas = np.asarray(betas, dtype=np.float64)
    rot = np.asarray(rot, dtype=np.float64)
    V = vt.shape[0]
    # eq. 8: shape blend shapes, vertex by vertex
    v_shaped = np.empty((V, 3))
    for v in range(V):
        v_shaped[v] = vt[v] + sum(betas[l] * S[v, :, l] for l in range(10))
    # eq. 10: joint locations regressed from the shaped (unposed) vertices
    joints = np.zeros((24, 3))
    for k in range(24):
        nz = np.nonzero(Jr[k])[0]
        for v in nz:
            joints[k] += Jr[k, v] * v_shaped[v]
    # eq. 9: pose blend shapes, linear in the elements of (R_k - I), k = 1..23, row-major per joint
    coeff = np.concatenate([(rot[k] - np.eye(3)).reshape(9) for k in range(1, 24)])      # [207]
    v_posed = np.empty((V, 3))
    for v in range(V):
        v_posed[v] = v_shaped[v] + coeff @ P[:, 3 * v:3 * v + 3]
    # eqs. 3-4: world transforms in the posed and in the rest configuration
    G = _chain(rot, joints, parents)
    G_rest = _chain(np.tile(np.eye(3), (24, 1, 1)), joints, parents)
    G_rel = [G[k] @ np.linalg.inv(G_rest[k]) for k in range(24)]
    # eq. 2: linear blend skinning, vertex by vertex
    verts = np.empty((V, 3))
    for v in range(V):
        acc = np.zeros(4)
        hv = np.array([v_posed[v, 0], v_posed[v, 1], v_posed[v, 2], 1.0])
        for k in np.nonzero(W[v])[0]:
            acc += W[v, k] * (G_rel[k] @ hv)
        verts[v] = acc[:3]
    posed_joints = np.stack([G[k][:3, 3] for k in range(24)])
    return verts, posed_joints


def joints49_one(model: dict, verts, posed_joints):
    """scripts/smpl.py:75-78 for one pose: 24 posed joints + 21 vertex picks + 9 extra-regressor joints,
    gathered by joint_map."""
    picks = np.asarray(model["vertex_picks"])
    extra = np.asarray(model["J_regressor_extra"], dtype=np.float64)
    stack = [posed_joints[k] for k in range(24)] + [verts[int(i)] for i in picks]
    for e in range(extra.shape[0]):
        stack.append(sum(extra[e, v] * verts[v] for v in np.nonzero(extra[e])[0]))
    stack = np.stack(stack)
    return stack[np.asarray(model["joint_map"])]
