"""Silhouette term: csrc/jrr_silhouette.cu + jrr_b200.mesh_renderer against oracle/silhouette_oracle.py
(pytorch3d 0.3.0 restated: PARITY UNPINNED).  CPU tests pin the oracle on analytic cases and the host-side helpers."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import silhouette_oracle as so  # noqa: E402

DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------ oracle known answers (CPU)
def test_oracle_single_triangle_known_answers():
    """One triangle facing the camera: coverage = the pixel centres strictly inside it, alpha = sigmoid(d^2 / sigma) with d
    the distance to the closest edge, zero background; a triangle behind the camera is culled."""
    S = 16
    f = 5000.0 / S
    Z = 100.0
    # ready mesh (flip_scale=False): ndc = f * (x, y) / Z.  Right triangle with ndc corners (-0.5,-0.5), (0.6,-0.5), (-0.5,0.6)
    ndc = torch.tensor([[-0.5, -0.5], [0.6, -0.5], [-0.5, 0.6]], dtype=torch.double)
    verts = torch.cat([ndc * Z / f, torch.zeros(3, 1, dtype=torch.double)], dim=1)[None]
    cam = torch.tensor([[0.0, 0.0, Z]], dtype=torch.double)
    faces = torch.tensor([[0, 1, 2]])
    alpha, p2f = so.soft_silhouette(verts, cam, faces, S, flip_scale=False)
    px, py = so.pixel_centres(S, torch.double)
    inside = (px > -0.5) & (py > -0.5) & (px + py < 0.1)
    assert torch.equal((p2f.reshape(-1) >= 0), inside)
    d = torch.minimum(torch.minimum(px + 0.5, py + 0.5), (0.1 - px - py) / 2 ** 0.5)
    expect = torch.where(inside, torch.sigmoid(d ** 2 / 1e-4), torch.zeros_like(d))
    assert (alpha.reshape(-1) - expect).abs().max().item() < 1e-12
    # row 0 is the TOP of the image (+Y up) and column 0 the side of +X (pytorch3d's +X-left convention)
    assert py[0] > py[-1] and px[0] > px[S - 1]
    behind, p2f_b = so.soft_silhouette(verts, -cam, faces, S, flip_scale=False)
    assert behind.abs().max().item() == 0 and (p2f_b < 0).all()


def test_oracle_nearest_face_wins_and_gradient_moves_the_edge():
    S = 16
    f = 5000.0 / S
    # (an edge 0.0025 ndc from a column of pixel centres: farther than ~0.06 the sigmoid saturates to exactly 1 in fp64)
    tri = torch.tensor([[-0.44, -0.5], [0.6, -0.5], [-0.44, 0.6]], dtype=torch.double)
    near = torch.cat([tri * 90.0 / f, torch.full((3, 1), -10.0, dtype=torch.double)], dim=1)      # view z = 90
    far = torch.cat([tri * 100.0 / f, torch.zeros(3, 1, dtype=torch.double)], dim=1)              # view z = 100
    verts = torch.cat([far, near])[None].clone().requires_grad_(True)
    cam = torch.tensor([[0.0, 0.0, 100.0]], dtype=torch.double)
    faces = torch.tensor([[0, 1, 2], [3, 4, 5]])
    alpha, p2f = so.soft_silhouette(verts, cam, faces, S, flip_scale=False)
    assert set(p2f.unique().tolist()) == {-1, 1}                   # the nearer copy hides the farther one everywhere
    alpha.sum().backward()
    g = verts.grad[0]
    assert g[:3].abs().max().item() == 0 and g[3:, :2].abs().max().item() > 0 and torch.isfinite(g).all()


def test_obj_face_reader_and_vertex_face_csr(tmp_path, jrr):
    p = tmp_path / "m.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvt 0 0\nf 1/1 2/1 3/1\nf 2//1 4//1 3//1 1//1\n")
    faces = jrr.load_obj_faces(str(p))
    assert faces.tolist() == [[0, 1, 2], [1, 3, 2], [1, 2, 0]]
    vertex_face_csr = jrr.mesh_renderer.vertex_face_csr
    f32, ptr, idx = vertex_face_csr(faces, 4)
    assert ptr.tolist() == [0, 2, 5, 8, 9]                       # vertex 0 in 2 corners, 1 in 3, 2 in 3, 3 in 1
    for v in range(4):
        ent = idx[ptr[v]:ptr[v + 1]].tolist()
        assert ent == sorted(ent) and all(f32.reshape(-1)[e] == v for e in ent)
    with pytest.raises(jrr.JrrError):
        vertex_face_csr([[0, 1, 4]], 4)
    local = jrr.synthetic.make_local_faces(jrr.synthetic.make_smpl_model(0)["v_template"])
    assert local.shape == (13776, 3) and local.min() >= 0 and local.max() < 6890
    assert (local[:, 0] != local[:, 1]).all() and (local[:, 1] != local[:, 2]).all()


# ------------------------------------------------------------------ CUDA vs oracle
def _scene(jrr, B, S, seed=3):
    model = jrr.synthetic.make_smpl_model(0)
    smpl = jrr.SMPL(model_dict=model, create_transl=False).to(DEV)
    inp = jrr.synthetic.make_pose_inputs(B, seed)
    R = torch.from_numpy(inp["true_rotmat"]).to(DEV)
    betas = torch.from_numpy(inp["true_betas"]).to(DEV)
    faces = jrr.synthetic.make_local_faces(model["v_template"], lbs_weights=model["lbs_weights"])
    g = torch.Generator().manual_seed(seed)
    cam = torch.tensor([0.0, 0.4, 5000.0 / S * 2.3]) + torch.randn(B, 3, generator=g) * torch.tensor([0.1, 0.1, 1.0])
    return smpl, R, betas, faces, cam.to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("S", [48, 96])
def test_silhouette_forward_and_backward_match_oracle(S, jrr):
    B = 2
    smpl, R, betas, faces, cam = _scene(jrr, B, S)
    rend = jrr.Mesh_Renderer(image_size=S, faces=faces)
    verts = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False).vertices.detach()
    v_req, c_req = verts.clone().requires_grad_(True), cam.clone().requires_grad_(True)
    alpha = rend.silhouette(c_req, v_req, flip_scale=True)
    # the oracle in fp64 on the same fp32 vertices
    vo, co = verts.double().cpu().requires_grad_(True), cam.double().cpu().requires_grad_(True)
    fo = torch.from_numpy(faces)
    alpha_o, p2f_o = so.soft_silhouette(vo, co, fo, S)
    mesh = rend.mesh(6890, DEV)
    _forward = jrr.mesh_renderer._forward
    _, p2f, _ = _forward(mesh, verts, cam, S, True)
    covered = (p2f_o >= 0).sum().item()
    mism = (p2f.cpu().long() != p2f_o).sum().item()
    print(f"[silhouette S={S}] covered {covered} of {B * S * S} pixels, winner-face mismatches {mism}")
    assert covered > 0.03 * B * S * S
    # every disagreement must be a round-off tie: the two faces cover the pixel at depths equal to 1e-5 (overlapping faces of
    # the triangle soup), or the pixel centre sits within 1e-5 (barycentric) of an edge of the face that one side rejects
    assert mism <= 0.05 * covered
    x, y, z = so.project(verts.double().cpu(), cam.double().cpu(), S)
    px, py = so.pixel_centres(S, torch.double)

    def bary_depth(b, f, k):
        i0, i1, i2 = fo[f]
        a = so._edge(x[b, i2], y[b, i2], x[b, i0], y[b, i0], x[b, i1], y[b, i1]) + so.EPS
        w = torch.stack([so._edge(px[k], py[k], x[b, i1], y[b, i1], x[b, i2], y[b, i2]),
                         so._edge(px[k], py[k], x[b, i2], y[b, i2], x[b, i0], y[b, i0]),
                         so._edge(px[k], py[k], x[b, i0], y[b, i0], x[b, i1], y[b, i1])]) / a
        return w.min().item(), (w * torch.stack([z[b, i0], z[b, i1], z[b, i2]])).sum().item()

    for b, r, c in torch.nonzero(p2f.cpu().long() != p2f_o).tolist():
        k = r * S + c
        fc, fr_ = int(p2f[b, r, c]), int(p2f_o[b, r, c])
        wc, zc = bary_depth(b, fc, k) if fc >= 0 else (0.0, float("inf"))
        wr, zr = bary_depth(b, fr_, k) if fr_ >= 0 else (0.0, float("inf"))
        tie = abs(zc - zr) <= 1e-5 * abs(zr) and wc > -1e-5 and wr > -1e-5
        on_edge = abs(wc) < 1e-5 or abs(wr) < 1e-5
        assert tie or on_edge, (b, r, c, fc, fr_, wc, zc, wr, zr)
    same = (p2f.cpu().long() == p2f_o)
    assert (alpha.detach().cpu().double() - alpha_o.detach())[same].abs().max().item() < 1e-4
    assert alpha.detach()[p2f < 0].abs().max().item() == 0
    # backward on IDENTICAL coverage (the CUDA winner map handed to the oracle), random upstream gradient
    g = torch.Generator().manual_seed(1)
    up = torch.randn(B, S, S, generator=g)
    alpha.backward(up.to(DEV))
    alpha_o2, _ = so.soft_silhouette(vo, co, fo, S, pix_to_face=p2f.cpu())
    alpha_o2.backward(up.double())
    ev, ec = rel(v_req.grad, vo.grad), rel(c_req.grad, co.grad)
    print(f"[silhouette S={S}] d/d vertices rel err {ev:.2e}, d/d cam rel err {ec:.2e}")
    assert ev < 1e-3 and ec < 1e-3
    # fixed-order reductions: a second run is bit-identical
    v2, c2 = verts.clone().requires_grad_(True), cam.clone().requires_grad_(True)
    a2 = rend.silhouette(c2, v2, flip_scale=True)
    a2.backward(up.to(DEV))
    assert torch.equal(a2, alpha) and torch.equal(v2.grad, v_req.grad) and torch.equal(c2.grad, c_req.grad)


@pytest.mark.gpu
def test_silhouette_fused_mse_and_reference_call_shapes(jrr):
    """silhouette_mse (loss + gradient seed inside the kernels) vs the oracle's optimize.py:234-236; Mesh_Renderer.forward
    returns the reference's [B,4,S,S]; render_mesh back-propagates into the body model's parameters."""
    B, S = 3, 48
    smpl, R, betas, faces, cam = _scene(jrr, B, S, seed=5)
    rend = jrr.Mesh_Renderer(image_size=S, faces=faces)
    verts = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False).vertices.detach()
    g = torch.Generator().manual_seed(2)
    target = (torch.rand(B, 1, S, S, generator=g) > 0.5).float().to(DEV)
    loss, dverts, dcam, alpha = jrr.silhouette_mse(rend, verts, cam, target, logical_batch=2 * B, weight=100.0)
    _forward = jrr.mesh_renderer._forward
    _, p2f, _ = _forward(rend.mesh(6890, DEV), verts, cam, S, True)
    vo, co = verts.double().cpu().requires_grad_(True), cam.double().cpu().requires_grad_(True)
    lo, _, _ = so.silhouette_loss(vo, co, torch.from_numpy(faces), target[:, 0].double().cpu(), S, logical_batch=2 * B,
                                  pix_to_face=p2f.cpu())
    (100.0 * lo).backward()
    el, ev, ec = abs(loss.item() - lo.item()) / lo.item(), rel(dverts, vo.grad), rel(dcam, co.grad)
    print(f"[silhouette mse] loss {loss.item():.6f} rel err {el:.2e}, d/d vertices {ev:.2e}, d/d cam {ec:.2e}")
    assert el < 1e-5 and ev < 1e-3 and ec < 1e-3
    # the reference's call shapes
    flipped = verts * torch.tensor([-2.0, -2.0, 2.0], device=DEV)
    img = rend({"cam": cam}, flipped)
    assert img.shape == (B, 4, S, S) and torch.equal(img[:, 3], alpha) and (img[:, :3] == 1).all()
    b_req = betas.clone().requires_grad_(True)
    R_req = R.clone().requires_grad_(True)
    out = jrr.render_mesh(smpl, rend, b_req, R_req[:, :1], R_req[:, 1:], {"cam": cam})
    assert out.shape == (B, 1, S, S) and torch.equal(out[:, 0], alpha)
    torch.nn.functional.mse_loss(out, target).backward()
    assert torch.isfinite(b_req.grad).all() and b_req.grad.abs().max().item() > 0
    assert torch.isfinite(R_req.grad).all() and R_req.grad.abs().max().item() > 0


@pytest.mark.gpu
def test_silhouette_edge_cases(jrr):
    """Mesh behind the camera -> empty image and zero gradients; degenerate faces and a face spanning the whole image do not
    produce NaNs; bad arguments surface as errors."""
    S = 32
    rend = jrr.Mesh_Renderer(image_size=S, faces=np.array([[0, 1, 2], [1, 1, 2], [3, 3, 3]]))
    f = 5000.0 / S
    tri = torch.tensor([[-3.0, -3.0], [3.0, -3.0], [0.0, 3.0], [0.1, 0.1]]) * 100.0 / f
    verts = torch.cat([tri, torch.zeros(4, 1)], dim=1)[None].to(DEV).requires_grad_(True)
    cam = torch.tensor([[0.0, 0.0, 100.0]], device=DEV)
    a = rend.silhouette(cam, verts, flip_scale=False)
    assert (a > 0.5).all()                                      # the big triangle covers every pixel centre
    a.sum().backward()
    assert torch.isfinite(verts.grad).all()
    # a box of several hundred pixel centres is walked by the whole warp (SIL_BIG): same image and gradient as the oracle
    # (triangle inside the image, so that pixel centres sit next to its edges -- far from an edge the sigmoid saturates)
    tri2 = torch.tensor([[-0.8, -0.7], [0.7, -0.6], [0.1, 0.8], [0.1, 0.1]]) * 100.0 / f
    v2 = torch.cat([tri2, torch.zeros(4, 1)], dim=1)[None].to(DEV).requires_grad_(True)
    g = torch.Generator().manual_seed(0)
    up = torch.randn(1, S, S, generator=g)
    a2 = rend.silhouette(cam, v2, flip_scale=False)
    (a2 * up.to(DEV)).sum().backward()
    vo = v2.detach().double().cpu().requires_grad_(True)
    ao, p2fo = so.soft_silhouette(vo, cam.double().cpu(), torch.tensor([[0, 1, 2], [1, 1, 2], [3, 3, 3]]), S, flip_scale=False)
    assert 200 < (p2fo == 0).sum().item() < S * S and (a2.detach().cpu().double() - ao.detach()).abs().max().item() < 1e-4
    (ao * up.double()).sum().backward()
    assert vo.grad.abs().max().item() > 1e-2 and rel(v2.grad, vo.grad) < 2e-3, rel(v2.grad, vo.grad)
    b = rend.silhouette(-cam, verts.detach(), flip_scale=False)
    assert b.abs().max().item() == 0
    with pytest.raises(jrr.JrrError):
        jrr.Mesh_Renderer(image_size=S, faces=np.array([[0, 1, 7]])).silhouette(cam, verts.detach())
    with pytest.raises(jrr.JrrError):
        rend.silhouette(cam.cpu(), verts.detach().cpu())


@pytest.fixture(scope="module")
def smpl_tc(jrr, model):
    return jrr.SMPL(model_dict=model, create_transl=False, gemm_impl=0).to(DEV)


@pytest.mark.gpu
def test_refine_with_silhouette_term_matches_oracle(smpl_tc, jrr, oracle, osmpl32, osmpl64, critic_sd, J_shipped):
    """Two iterations of optimize.py:220-265 with ALL the reference's terms -- 3-D joints, pose critic, 2-D reprojection and
    the silhouette (x100) -- through PoseRefiner.refine_silhouette (rasteriser + module backward feeding the fused step as an
    external gradient) vs the oracle composition; image size 224 as in optimize.py:111."""
    from test_gpu_parity import _cam_problem
    B, S = 4, 224
    fr, gt2d, cam0 = _cam_problem(jrr, oracle, osmpl32, J_shipped, B, 11)
    faces = jrr.synthetic.make_local_faces(smpl_tc._model_np["v_template"], lbs_weights=smpl_tc._model_np["lbs_weights"])
    rend = jrr.Mesh_Renderer(image_size=S, faces=faces)
    g = torch.Generator().manual_seed(3)
    mask = (torch.rand(B, 1, S, S, generator=g) > 0.7).float()
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    x6, be, cam = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone(), cam0.to(DEV).clone()
    # coverage of the FIRST iteration as the CUDA rasteriser sees it (handed to the oracle: identical winner maps)
    _forward = jrr.mesh_renderer._forward
    R0 = jrr.rot6d_to_rotmat(x6.reshape(-1, 6)).reshape(B, 24, 3, 3)
    v0 = smpl_tc(betas=be, body_pose=R0[:, 1:], global_orient=R0[:, :1], pose2rot=False).vertices
    _, p2f, _ = _forward(rend.mesh(6890, DEV), v0.contiguous(), cam, S, True)
    assert (p2f >= 0).float().mean().item() > 0.02
    sil = dict(faces=torch.from_numpy(faces), target=mask[:, 0], S=S, weight=100.0, pix_to_face=p2f.cpu())
    x6o, bo, co, hist = oracle.refine_2d(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], cam0, fr["gt_mm"], gt2d,
                                         iters=1, silhouette=sil)
    x6p, bp, cp, _ = oracle.refine_2d(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], cam0, fr["gt_mm"], gt2d, iters=1)
    loss, loss_s = ref.refine_silhouette(x6, be, cam, fr["gt_mm"].to(DEV), gt2d.to(DEV), mask.to(DEV), rend, iters=1)
    torch.cuda.synchronize()
    d6, db, dc = (x6.cpu() - x6o).abs().max().item(), (be.cpu() - bo).abs().max().item(), (cam.cpu() - co).abs().max().item()
    moved = (x6o - x6p).abs().max().item()
    print(f"refine + silhouette, 1 step: |dx6| {d6:.2e} |dbetas| {db:.2e} |dcam| {dc:.2e}; the term moves the step by {moved:.2e}; "
          f"silhouette loss {loss_s.item():.5f}")
    # Adam's first step is lr * sign(g): parameters whose two gradient terms nearly cancel may flip (2 lr apart); everything
    # else must agree closely
    close = ((x6.cpu() - x6o).abs() < 2e-4).float().mean().item()
    assert close > 0.995 and db < 2.1e-2 and dc < 2.1e-2, (close, db, dc)
    assert moved > 1e-3                                   # the silhouette term changes the update (it is not a no-op)
    assert ref.native._ext == (None, None, None)          # the external gradient is cleared afterwards
    # the gradient itself: Adam's first moment after one step is 0.1 g -- against autograd of the oracle's total loss
    # (fp64 oracle: the camera gradient is a sum over 6890 vertices that cancels to a few percent of its terms)
    sd64 = {k: v.double() for k, v in critic_sd.items()}
    xr, br, cr = (t.double().clone().requires_grad_(True) for t in (fr["x6"], fr["betas"], cam0))
    total, _, _, pred = oracle.refine_loss(osmpl64, J_shipped.double(), sd64, xr, br, fr["gt_mm"].double(), 10000.0, 10.0, None)
    total = total + 0.01 * ((gt2d.double() - oracle.project_2d(pred, cr)) ** 2).sum() / (B * 17 * 2)
    Rr = oracle.rot6d_to_rotmat(xr.reshape(-1, 6)).view(B, 24, 3, 3)
    vr = osmpl64(betas=br, body_pose=Rr[:, 1:], global_orient=Rr[:, :1], pose2rot=False).vertices
    ls, _, _ = so.silhouette_loss(vr, cr, sil["faces"], sil["target"].double(), S, pix_to_face=sil["pix_to_face"])
    (total + 100.0 * ls).backward()
    st = ref._buffers(B, two_d=True)
    g_cuda = st["m"].cpu() / 0.1
    g_or = torch.cat([xr.grad.reshape(B, 144), br.grad], dim=1)
    eg, ecam = rel(g_cuda, g_or), rel(st["cm"].cpu() / 0.1, cr.grad)
    print(f"refine + silhouette: parameter gradient rel err {eg:.2e}, camera gradient rel err {ecam:.2e}, "
          f"silhouette loss {loss_s.item():.6f} vs {ls.item():.6f}")
    assert eg < 1e-3 and ecam < 5e-3
    assert abs(loss_s.item() - ls.item()) / ls.item() < 1e-4
    # and plain refine_2d afterwards is unaffected by the (cleared) hook
    x6b, beb, camb = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone(), cam0.to(DEV).clone()
    ref.refine_2d(x6b, beb, camb, fr["gt_mm"].to(DEV), gt2d.to(DEV), iters=1)
    assert (x6b.cpu() - x6p).abs().max().item() < 2e-4


@pytest.mark.gpu
def test_refinement_loop_with_silhouette_masks(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped):
    """RefinementLoop.run_batch with 'mask_rcnn' in the batch and a renderer: camera fit, refinement with all five terms,
    critic step, refit.  Masks rendered from the TRUE poses: the silhouette loss must fall over the iterations, and the
    regressor / critic updates stay finite."""
    from test_gpu_parity import _cam_problem
    B, S = 6, 64
    fr, gt2d, cam0 = _cam_problem(jrr, oracle, osmpl32, J_shipped, B, 21)
    # the 2-D joints of _cam_problem live on the 224-pixel screen of renderer.py; the silhouettes are rendered at S
    faces = jrr.synthetic.make_local_faces(smpl_tc._model_np["v_template"], lbs_weights=smpl_tc._model_np["lbs_weights"])
    rend = jrr.Mesh_Renderer(image_size=S, faces=faces)
    Rt = fr["true_rotmat"].to(DEV)
    cam_sil = torch.tensor([0.0, 0.4, 5000.0 / S * 2.3], device=DEV).repeat(B, 1)
    with torch.no_grad():
        true_img = jrr.render_mesh(smpl_tc, rend, fr["true_betas"].to(DEV), Rt[:, :1], Rt[:, 1:], {"cam": cam_sil})
    mask = (true_img > 0.5).float()
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    x6, be, cam = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone(), cam_sil.clone()
    gt, g2 = fr["gt_mm"].to(DEV), torch.full((B, 17, 2), 112.0, device=DEV)
    _, s0 = ref.refine_silhouette(x6, be, cam, gt, g2, mask, rend, iters=1, w_2d=0.0)
    s0 = s0.item()
    _, s1 = ref.refine_silhouette(x6, be, cam, gt, g2, mask, rend, iters=30, w_2d=0.0)
    print(f"silhouette loss against masks of the true poses: {s0:.5f} -> {s1.item():.5f} after 30 iterations")
    assert s1.item() < s0
    # graph replay of the iteration (first eager, second captured, rest replayed) = eager launches, bit for bit
    res = []
    for ug in (False, True):
        xa, ba, ca = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone(), cam_sil.clone()
        la, sa = ref.refine_silhouette(xa, ba, ca, gt, g2, mask, rend, iters=5, w_2d=0.0, use_graph=ug)
        res.append((xa, ba, ca, la.clone(), sa.clone()))
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    loop = jrr.RefinementLoop(smpl_tc, J_shipped, critic_sd, refine_iters=3, cam_iters=5, silhouette_renderer=rend)
    out = loop.run_batch({"orient": fr["x6"][:, :1], "pose": fr["x6"][:, 1:], "betas": fr["betas"], "gt_j3d": fr["gt_mm"],
                          "gt_j2d": gt2d, "cam": cam0, "mask_rcnn": mask})
    assert "silhouette_loss" in out and torch.isfinite(out["silhouette_loss"]).all()
    assert torch.isfinite(out["refine_loss"]).all() and torch.isfinite(out["refit_loss"]).all()
    assert torch.isfinite(out["x6"]).all() and torch.isfinite(out["cam"]).all()
