"""CPU restatement of the reference's soft-silhouette renderer -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file; the product
package never does.

What it restates (PARITY UNPINNED: pytorch3d==0.3.0 is a dependency of the reference that is neither vendored
under /root/reference nor installable offline; there are no reference tests or golden images for this path):

* ``Mesh_Renderer.forward`` -- ``/root/reference/scripts/mesh_renderer.py:23-79``:
  ``PerspectiveCameras(T=cam, focal_length=5000/image_size, principal_point=0)`` (R = I),
  ``MeshRasterizer(RasterizationSettings(image_size, blur_radius=0.0, faces_per_pixel=1))``,
  ``SoftSilhouetteShader(BlendParams(sigma=1e-4, gamma=1e-4))``; the output image is ``[B,4,S,S]`` and the loss
  uses channel 3.
* ``render_mesh`` -- ``scripts/optimize.py:77-85``: vertices ``* (-2, -2, 2)`` before rendering.
* the silhouette loss -- ``optimize.py:234-236``: ``nn.MSELoss()(img, batch['mask_rcnn'])``.

pytorch3d 0.3.0's published algorithm, as restated here (its ``rasterize_meshes.cu`` / ``geometry_utils.cuh`` /
``blending.py``):

* transform: ``view = X + T``; ``ndc.xy = f * view.xy / view.z``; the rasteriser keeps ``z = view.z``.
* pixel (row r, col c) has the centre ``x = -1 + (2 (S-1-c) + 1) / S``, ``y = -1 + (2 (S-1-r) + 1) / S``
  (+X left, +Y up).
* a face is skipped when ``zmax < 0`` or ``|EdgeFunction(v0, v1, v2)| <= 1e-8``; barycentric coordinates are
  ``w_i = EdgeFunction(p, v_j, v_k) / (EdgeFunction(v2, v0, v1) + 1e-8)`` (no perspective correction, no clipping);
  a pixel is covered when all three are ``> 0`` (blur_radius = 0 keeps inside pixels only) and the interpolated
  depth ``w . z >= 0``; with ``faces_per_pixel = 1`` the face with the smallest depth wins.
* ``dists`` of a covered pixel = ``-min`` over the three edges of the squared point-segment distance
  (segment parameter clamped to [0, 1]; an edge shorter than ``1e-8`` squared counts as its end point).
* ``alpha = 1 - prod_k (1 - sigmoid(-dists_k / sigma) * mask_k)`` = ``sigmoid(d2 / sigma)`` for one face per pixel,
  0 on the background.
* backward: only ``dists`` carries gradient to the face's projected corners; pytorch3d differentiates the
  point-segment distance with the segment parameter held fixed, which is what autograd gives here too (the
  distance is stationary in the parameter when it is not clamped, constant in it when it is).
"""
from __future__ import annotations

import torch

EPS = 1e-8


def project(verts, cam, image_size, flip_scale=True):
    """[B,V,3] vertices, [B,3] camera translation -> ndc x, y and view z, each [B,V]."""
    s = verts.new_tensor([-2.0, -2.0, 2.0]) if flip_scale else verts.new_ones(3)
    view = verts * s + cam[:, None, :]
    f = 5000.0 / image_size
    return f * view[..., 0] / view[..., 2], f * view[..., 1] / view[..., 2], view[..., 2]


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def pixel_centres(S, dtype):
    i = torch.arange(S, dtype=dtype)
    c = -1.0 + (2.0 * (S - 1 - i) + 1.0) / S         # index along the image axis -> ndc
    py, px = torch.meshgrid(c, c, indexing="ij")       # row r -> y, col c -> x
    return px.reshape(-1), py.reshape(-1)


@torch.no_grad()
def rasterize(x, y, z, faces, S, rows=4):
    """winner face per pixel of ONE frame (x, y, z: [V]); -1 = background.  Bands of `rows` image rows against the
    faces whose y range reaches the band (a pure speed-up: every face a pixel can be inside of is kept)."""
    x0, x1, x2 = x[faces[:, 0]], x[faces[:, 1]], x[faces[:, 2]]
    y0, y1, y2 = y[faces[:, 0]], y[faces[:, 1]], y[faces[:, 2]]
    z0, z1, z2 = z[faces[:, 0]], z[faces[:, 1]], z[faces[:, 2]]
    zmax = torch.maximum(z0, torch.maximum(z1, z2))
    farea = _edge(x0, y0, x1, y1, x2, y2)
    live = ~(zmax < 0) & ~((farea <= EPS) & (farea >= -EPS))
    area = _edge(x2, y2, x0, y0, x1, y1) + EPS
    ymin = torch.minimum(y0, torch.minimum(y1, y2))
    ymax = torch.maximum(y0, torch.maximum(y1, y2))
    px, py = pixel_centres(S, x.dtype)
    out = torch.full((S * S,), -1, dtype=torch.long)
    for r0 in range(0, S, rows):
        s, e = r0 * S, min(S, r0 + rows) * S
        band = py[s:e]
        sel = torch.nonzero(live & (ymax >= band.min()) & (ymin <= band.max())).reshape(-1)
        if sel.numel() == 0:
            continue
        qx, qy = px[s:e, None], band[:, None]
        w0 = _edge(qx, qy, x1[sel], y1[sel], x2[sel], y2[sel]) / area[sel]
        w1 = _edge(qx, qy, x2[sel], y2[sel], x0[sel], y0[sel]) / area[sel]
        w2 = _edge(qx, qy, x0[sel], y0[sel], x1[sel], y1[sel]) / area[sel]
        pz = w0 * z0[sel] + w1 * z1[sel] + w2 * z2[sel]
        ok = (w0 > 0) & (w1 > 0) & (w2 > 0) & ~(pz < 0)
        pz = torch.where(ok, pz, torch.full_like(pz, float("inf")))
        zmin, k = pz.min(dim=1)                          # (first index among equal depths = the lowest face id)
        out[s:e] = torch.where(torch.isfinite(zmin), sel[k], torch.full_like(k, -1))
    return out.reshape(S, S)


def _seg_dist(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    safe = torch.where(l2 <= EPS, torch.ones_like(l2), l2)
    t = ((bax * (px - ax) + bay * (py - ay)) / safe).clamp(0.0, 1.0)
    qx, qy = ax + t * bax, ay + t * bay
    d = (px - qx) ** 2 + (py - qy) ** 2
    return torch.where(l2 <= EPS, (px - bx) ** 2 + (py - by) ** 2, d)


def soft_silhouette(verts, cam, faces, S, sigma=1e-4, flip_scale=True, pix_to_face=None):
    """alpha [B,S,S] (differentiable w.r.t. verts and cam) and the winner map [B,S,S].  ``pix_to_face`` given: use that
    assignment instead of rasterising (to compare the differentiable half on identical coverage)."""
    B = verts.shape[0]
    x, y, z = project(verts, cam, S, flip_scale)
    px, py = pixel_centres(S, verts.dtype)
    alphas, maps = [], []
    for b in range(B):
        p2f = rasterize(x[b].detach(), y[b].detach(), z[b].detach(), faces, S) if pix_to_face is None else pix_to_face[b].long()
        maps.append(p2f)
        flat = p2f.reshape(-1)
        cov = flat >= 0
        f = faces[flat.clamp(min=0)]
        ax, ay = x[b][f[:, 0]], y[b][f[:, 0]]
        bx, by = x[b][f[:, 1]], y[b][f[:, 1]]
        cx, cy = x[b][f[:, 2]], y[b][f[:, 2]]
        d2 = torch.minimum(_seg_dist(px, py, ax, ay, bx, by),
                           torch.minimum(_seg_dist(px, py, ax, ay, cx, cy), _seg_dist(px, py, bx, by, cx, cy)))
        a = torch.sigmoid(d2 / sigma)                    # = sigmoid(-dists / sigma), dists = -d2 inside the face
        alphas.append(torch.where(cov, a, torch.zeros_like(a)).reshape(S, S))
    return torch.stack(alphas), torch.stack(maps)


def silhouette_loss(verts, cam, faces, target, S, sigma=1e-4, logical_batch=None, pix_to_face=None):
    """optimize.py:234-236 on the body model's vertices: MSE(render_mesh(...), mask) over [B_logical,1,S,S]."""
    alpha, p2f = soft_silhouette(verts, cam, faces, S, sigma, True, pix_to_face)
    LB = verts.shape[0] if logical_batch is None else logical_batch
    return ((alpha - target) ** 2).sum() / (LB * S * S), alpha, p2f
