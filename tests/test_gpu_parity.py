"""Parity of the CUDA path (through the C ABI) against the CPU oracle, on the B200.
Tolerances are the north star's: vertices/joints 1e-5 relative (max-abs difference over
max-abs reference), refined MPJPE within 0.01 mm, regressor weights within 1e-4."""
import numpy as np
import pytest
import torch

from conftest import make_frames

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def smpl_tc(jrr, model):
    return jrr.SMPL(model_dict=model, create_transl=False, gemm_impl=0).to(DEV)


@pytest.fixture(scope="module")
def smpl_simt(jrr, model):
    return jrr.SMPL(model_dict=model, create_transl=False, gemm_impl=1).to(DEV)


# ------------------------------------------------------------------ kernel level: the GEMM
@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 256, 224), (128, 224, 3456), (384, 1024, 768)])
def test_gemm_simt_vs_torch(smpl_simt, shape):
    M, N, K = shape
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    C = smpl_simt.native().debug_gemm(A, B, impl=1)
    ref = A.double() @ B.double().t()
    assert rel(C, ref) < 2e-6


@pytest.mark.parametrize("shape", [(128, 128, 32), (128, 128, 224), (256, 256, 224), (128, 224, 3456),
                                   (384, 1024, 768), (1024, 20736, 224)])
def test_gemm_tcgen05_3xtf32_vs_fp64(smpl_tc, shape):
    M, N, K = shape
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    C = smpl_tc.native().debug_gemm(A, B, impl=0)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    # 3xTF32 keeps ~21 mantissa bits per product; the tensor core accumulates in fp32 with
    # truncation, so the error grows ~linearly with the number of K steps (K/8 per pass)
    err = rel(C, ref)
    print(f"tcgen05 3xTF32 {shape}: rel err {err:.2e}")
    assert err < (1e-5 if K <= 1024 else 4e-5)


@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 256, 224), (384, 1024, 768), (4096, 1024, 1024), (512, 768, 1024)])
def test_gemm_tcgen05_a_through_tmem_vs_fp64(smpl_tc, shape):
    """The variant that takes a plain fp32 A, splits it in registers and feeds the MMAs from tensor memory."""
    M, N, K = shape
    g = torch.Generator(device="cpu").manual_seed(M + N + K + 1)
    A = torch.randn(M, K, generator=g).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    C = smpl_tc.native().debug_gemm(A, B, impl=2)
    torch.cuda.synchronize()
    err = rel(C, A.double() @ B.double().t())
    print(f"tcgen05 3xTF32 A-through-TMEM {shape}: rel err {err:.2e}")
    assert err < 1e-5
    assert torch.equal(C, smpl_tc.native().debug_gemm(A, B, impl=2))


# ------------------------------------------------------------------ SMPL forward (config C1)
@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("pose2rot", [False, True])
def test_smpl_forward_matches_oracle(impl, pose2rot, smpl_tc, smpl_simt, osmpl32, osmpl64, jrr):
    smpl = smpl_tc if impl == "tc" else smpl_simt
    inp = jrr.synthetic.make_pose_inputs(64, 3)
    betas = torch.from_numpy(inp["true_betas"])
    if pose2rot:
        g = torch.Generator().manual_seed(5)
        go = torch.randn(64, 3, generator=g)
        bp = 0.3 * torch.randn(64, 69, generator=g)
        kw = dict(global_orient=go, body_pose=bp)
    else:
        R = torch.from_numpy(inp["true_rotmat"])
        kw = dict(global_orient=R[:, :1], body_pose=R[:, 1:])
    ref = osmpl64(betas=betas.double(), pose2rot=pose2rot, **{k: v.double() for k, v in kw.items()})
    ref32 = osmpl32(betas=betas, pose2rot=pose2rot, **kw)
    out = smpl(betas=betas.to(DEV), pose2rot=pose2rot, **{k: v.to(DEV) for k, v in kw.items()})
    assert out.vertices.shape == (64, 6890, 3) and out.joints.shape == (64, 49, 3)
    ev, ej = rel(out.vertices, ref.vertices), rel(out.joints, ref.joints)
    print(f"[{impl} pose2rot={pose2rot}] vertices rel {ev:.2e} joints rel {ej:.2e}; "
          f"fp32 oracle vs fp64: {rel(ref32.vertices, ref.vertices):.2e}")
    assert ev < 1e-5 and ej < 1e-5


def test_smpl_forward_known_answers(smpl_tc, model):
    """identity rotations + zero betas -> template; global rotation only -> rigid motion."""
    B = 3
    I = torch.eye(3, device=DEV).expand(B, 24, 3, 3).contiguous()
    z = torch.zeros(B, 10, device=DEV)
    out = smpl_tc(betas=z, body_pose=I[:, 1:], global_orient=I[:, :1], pose2rot=False)
    vt = torch.from_numpy(model["v_template"]).to(DEV)
    assert (out.vertices - vt[None]).abs().max().item() < 2e-6
    c, s = np.cos(0.7), np.sin(0.7)
    R0 = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], device=DEV, dtype=torch.float32)
    Rg = I.clone()
    Rg[:, 0] = R0
    out = smpl_tc(betas=z, body_pose=Rg[:, 1:], global_orient=Rg[:, :1], pose2rot=False)
    J0 = torch.from_numpy(model["J_regressor"][0] @ model["v_template"]).to(DEV)
    expect = (vt - J0) @ R0.t() + J0
    assert (out.vertices - expect[None]).abs().max().item() < 5e-6


def test_smpl_ragged_batch_and_beta_broadcast(smpl_tc, osmpl32, jrr):
    """B not a multiple of the 128-pose tile, batch-1 betas broadcast (smplx semantics)."""
    inp = jrr.synthetic.make_pose_inputs(131, 9)
    R = torch.from_numpy(inp["true_rotmat"])
    b1 = torch.from_numpy(inp["true_betas"][:1])
    ref = osmpl32(betas=b1, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    out = smpl_tc(betas=b1.to(DEV), body_pose=R[:, 1:].to(DEV), global_orient=R[:, :1].to(DEV), pose2rot=False)
    assert rel(out.vertices, ref.vertices) < 1e-5 and rel(out.joints, ref.joints) < 1e-5


@pytest.mark.parametrize("B", [1, 3, 4, 5, 8, 9])
def test_smpl_small_batch_single_launch_forward(B, smpl_tc, osmpl64, jrr):
    """Up to 8 poses SMPL.forward is ONE launch (warp per vertex, csrc/jrr_pose.cu smpl_small_fwd_kernel); 9 poses take
    the tensor-core path.  All three rotation formats, vertices + 49 joints vs the fp64 oracle, joints-only calls, and a
    second call (the device counter that elects the joint-gathering block resets itself)."""
    inp = jrr.synthetic.make_pose_inputs(64, 21)
    R = torch.from_numpy(inp["true_rotmat"][:B])
    betas = torch.from_numpy(inp["true_betas"][:B])
    nat = smpl_tc.native()
    ref = osmpl64(betas=betas.double(), body_pose=R[:, 1:].double(), global_orient=R[:, :1].double(), pose2rot=False)
    for rep_ in range(2):
        out = smpl_tc(betas=betas.to(DEV), body_pose=R[:, 1:].to(DEV), global_orient=R[:, :1].to(DEV), pose2rot=False)
        assert nat.launches == (1 if B <= 8 else 4), nat.launches
        ev, ej = rel(out.vertices, ref.vertices), rel(out.joints, ref.joints)
        assert ev < 1e-5 and ej < 1e-5, (ev, ej)
    print(f"[small forward B={B}] vertices rel {ev:.2e} joints rel {ej:.2e} launches {nat.launches}")
    # joints only (no vertex buffer from the caller), rot6d input
    x6 = torch.from_numpy(inp["x6"][:B])
    R6 = jrr.rot6d_to_rotmat(x6.reshape(-1, 6)).reshape(B, 24, 3, 3)
    ref6 = osmpl64(betas=betas.double(), body_pose=R6[:, 1:].double(), global_orient=R6[:, :1].double(), pose2rot=False)
    _, j6 = nat.smpl_forward(betas.to(DEV), x6.to(DEV).reshape(B, 24, 6), 2, False, True)
    assert rel(j6, ref6.joints) < 1e-5
    # axis-angle
    g = torch.Generator().manual_seed(5)
    go, bp = torch.randn(B, 3, generator=g), 0.3 * torch.randn(B, 69, generator=g)
    refa = osmpl64(betas=betas.double(), global_orient=go.double(), body_pose=bp.double(), pose2rot=True)
    outa = smpl_tc(betas=betas.to(DEV), global_orient=go.to(DEV), body_pose=bp.to(DEV), pose2rot=True)
    assert rel(outa.vertices, refa.vertices) < 1e-5 and rel(outa.joints, refa.joints) < 1e-5


# ------------------------------------------------------------------ SMPL backward
@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("pose2rot", [False, True])
def test_smpl_backward_matches_oracle_autograd(impl, pose2rot, smpl_tc, smpl_simt, osmpl64, jrr):
    smpl = smpl_tc if impl == "tc" else smpl_simt
    B = 48
    inp = jrr.synthetic.make_pose_inputs(B, 11)
    g = torch.Generator().manual_seed(17)
    betas = torch.from_numpy(inp["true_betas"])
    if pose2rot:
        go = torch.randn(B, 3, generator=g)
        bp = 0.3 * torch.randn(B, 69, generator=g)
    else:
        R = torch.from_numpy(inp["true_rotmat"])
        go, bp = R[:, :1].contiguous(), R[:, 1:].contiguous()
    wv = torch.randn(B, 6890, 3, generator=g)
    wj = torch.randn(B, 49, 3, generator=g)

    def run(fn, dt, dev):
        b = betas.to(dev, dt).requires_grad_(True)
        o = go.to(dev, dt).requires_grad_(True)
        p = bp.to(dev, dt).requires_grad_(True)
        out = fn(betas=b, body_pose=p, global_orient=o, pose2rot=pose2rot)
        loss = (out.vertices * wv.to(dev, dt)).sum() + (out.joints * wj.to(dev, dt)).sum()
        loss.backward()
        return b.grad, o.grad, p.grad

    rb, ro, rp = run(osmpl64, torch.float64, "cpu")
    gb, go_, gp = run(smpl, torch.float32, DEV)
    eb, eo, ep = rel(gb, rb), rel(go_, ro), rel(gp, rp)
    print(f"[{impl} pose2rot={pose2rot}] dbetas {eb:.2e} dorient {eo:.2e} dpose {ep:.2e}")
    assert eb < 1e-4 and eo < 1e-4 and ep < 1e-4


@pytest.mark.parametrize("B", [1, 3, 4, 5, 8, 9, 16])
@pytest.mark.parametrize("pose2rot", [False, True])
def test_smpl_small_batch_backward(B, pose2rot, smpl_tc, osmpl64, jrr):
    """Up to 16 poses (groups of 8) SMPL.backward is three / five launches (warp-per-vertex gradients, fixed-order reduction, chain backward:
    csrc/jrr_pose.cu smpl_small_bwd_kernel) instead of the padded tensor-core path: gradients w.r.t. betas / orientation /
    pose vs fp64 autograd, with vertex-only, joints-only and combined upstream gradients; reruns are bit-identical."""
    inp = jrr.synthetic.make_pose_inputs(16, 23)
    g = torch.Generator().manual_seed(29)
    betas = torch.from_numpy(inp["true_betas"][:B])
    if pose2rot:
        go, bp = torch.randn(B, 3, generator=g), 0.3 * torch.randn(B, 69, generator=g)
    else:
        R = torch.from_numpy(inp["true_rotmat"][:B])
        go, bp = R[:, :1].contiguous(), R[:, 1:].contiguous()
    wv, wj = torch.randn(B, 6890, 3, generator=g), torch.randn(B, 49, 3, generator=g)
    nat = smpl_tc.native()

    def run(fn, dt, dev, use_v, use_j):
        b = betas.to(dev, dt).requires_grad_(True)
        o = go.to(dev, dt).requires_grad_(True)
        p = bp.to(dev, dt).requires_grad_(True)
        out = fn(betas=b, body_pose=p, global_orient=o, pose2rot=pose2rot)
        loss = 0
        if use_v:
            loss = loss + (out.vertices * wv.to(dev, dt)).sum()
        if use_j:
            loss = loss + (out.joints * wj.to(dev, dt)).sum()
        loss.backward()
        return b.grad, o.grad, p.grad

    for use_v, use_j in ((True, True), (True, False), (False, True)):
        rb, ro, rp = run(osmpl64, torch.float64, "cpu", use_v, use_j)
        gb, go_, gp = run(smpl_tc, torch.float32, DEV, use_v, use_j)
        assert nat.launches == (3 if B <= 8 else 5), nat.launches      # groups of 8 poses: 2 launches each + the chain backward
        eb, eo, ep = rel(gb, rb), rel(go_, ro), rel(gp, rp)
        assert eb < 1e-4 and eo < 1e-4 and ep < 1e-4, (use_v, use_j, eb, eo, ep)
    gb2, go2, gp2 = run(smpl_tc, torch.float32, DEV, False, True)
    assert torch.equal(gb2, gb) and torch.equal(go2, go_) and torch.equal(gp2, gp)
    print(f"[small backward B={B} pose2rot={pose2rot}] dbetas {eb:.2e} dorient {eo:.2e} dpose {ep:.2e}, {nat.launches} launches")


# ------------------------------------------------------------------ find_joints / critic
@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_find_joints_matches_oracle(which, smpl_tc, osmpl64, oracle, jrr, J_shipped, J_dense, frames64):
    J = J_shipped if which == "shipped" else J_dense
    R = frames64["true_rotmat"]
    b = frames64["true_betas"]
    ref = oracle.find_joints(osmpl64, b.double(), R[:, :1].double(), R[:, 1:].double(), J.double(),
                             mask=oracle.find_j_reg_mask(J.double()))
    with torch.no_grad():
        out = jrr.find_joints(smpl_tc, b.to(DEV), R[:, :1].to(DEV), R[:, 1:].to(DEV), J.to(DEV),
                              mask=jrr.find_j_reg_mask(J.to(DEV)))
    assert out.shape == (64, 17, 3)
    assert rel(out, ref) < 1e-5
    # return_verts goes through the module path
    with torch.no_grad():
        out2, verts = jrr.find_joints(smpl_tc, b.to(DEV), R[:, :1].to(DEV), R[:, 1:].to(DEV), J.to(DEV),
                                      return_verts=True)
    assert rel(out2, ref) < 1e-5 and verts.shape == (64, 6890, 3)


@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_find_joints_autograd_through_fused_kernels(which, smpl_tc, osmpl64, oracle, jrr, J_shipped, J_dense, frames64):
    """utils.find_joints differentiated by the caller (renderer.py:27-28 -> the 2-D loss, optimize.py:193-199): forward and
    backward run on the fused loss-path kernels (jrr_find_joints / jrr_find_joints_backward, no torch contraction); gradients
    w.r.t. betas and the rotation matrices vs fp64 autograd; a regressor that requires grad still takes the eager path."""
    J = J_shipped if which == "shipped" else J_dense
    n = 70
    R, b = frames64["true_rotmat"][:n], frames64["true_betas"][:n]
    n = R.shape[0]
    g = torch.Generator().manual_seed(2)
    up = torch.randn(n, 17, 3, generator=g)
    Ro, bo = R.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = oracle.find_joints(osmpl64, bo, Ro[:, :1], Ro[:, 1:], J.double())
    (ref * up.double()).sum().backward()
    Rc, bc = R.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    nat = smpl_tc.native()
    out = jrr.find_joints(smpl_tc, bc, Rc[:, :1], Rc[:, 1:], J.to(DEV))
    assert isinstance(out.grad_fn, torch.autograd.function.BackwardCFunction) or "FindJoints" in type(out.grad_fn).__name__
    (out * up.to(DEV)).sum().backward()
    eb, er = rel(bc.grad, bo.grad), rel(Rc.grad, Ro.grad)
    print(f"[find_joints autograd {which}] joints {rel(out, ref):.2e}, d/d betas {eb:.2e}, d/d rotations {er:.2e}")
    assert rel(out, ref) < 1e-5 and eb < 1e-4 and er < 1e-4
    Jg = J.to(DEV).clone().requires_grad_(True)
    out2 = jrr.find_joints(smpl_tc, b.to(DEV), R[:, :1].to(DEV), R[:, 1:].to(DEV), Jg)
    out2.sum().backward()
    assert Jg.grad is not None and rel(out2, ref) < 1e-5


def test_find_joints_golden(smpl_tc, jrr, J_shipped):
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ref_utils_golden.npz"))
    R = torch.from_numpy(z["rotmat"]).to(DEV)
    with torch.no_grad():
        out = jrr.find_joints(smpl_tc, torch.from_numpy(z["betas"]).to(DEV), R[:, :1], R[:, 1:],
                              J_shipped.to(DEV))
    assert rel(out, torch.from_numpy(z["find_joints"])) < 1e-5
    mp, pa = jrr.evaluate(out, torch.from_numpy(z["gt_mm"]).to(DEV))
    assert abs(mp - float(z["mpjpe"])) < 1e-2 and abs(pa - float(z["pa_mpjpe"])) < 1e-2


def test_critic_forward_matches_oracle(smpl_tc, jrr, oracle, critic_sd):
    D = jrr.Discriminator()
    D.load_state_dict(critic_sd)           # reference state_dict layout loads unchanged
    D.bind(smpl_tc.native())
    g = torch.Generator().manual_seed(3)
    x6 = torch.randn(200, 24, 6, generator=g)
    ref = oracle.discriminator_forward({k: v.double() for k, v in critic_sd.items()}, x6.double())
    out = D(x6.to(DEV))
    assert out.shape == (200, 25, 1)
    assert (out.cpu().double() - ref).abs().max().item() < 2e-6
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ref_utils_golden.npz"))
    out = D(torch.from_numpy(z["x6"]).to(DEV))
    assert (out.cpu() - torch.from_numpy(z["critic_scores"])).abs().max().item() < 2e-6


# ------------------------------------------------------------------ the fused refinement step
def _oracle_refine(oracle, osmpl, J, critic_sd, fr, iters, **kw):
    return oracle.refine(osmpl, J, critic_sd, fr["x6"], fr["betas"], fr["gt_mm"], iters=iters, **kw)


@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_refine_single_step_gradients(which, smpl_tc, jrr, oracle, osmpl64, critic_sd, J_shipped, J_dense):
    """After ONE Adam step from zero state the update is -lr*g/(|g|+eps'): compare the
    implied per-parameter direction with the fp64 oracle's, and the losses."""
    J = J_shipped if which == "shipped" else J_dense
    fr = make_frames(jrr, oracle, oracle.OracleSMPL(jrr.synthetic.make_smpl_model(0)), J, 40, 21)
    sd64 = {k: v.double() for k, v in critic_sd.items()}
    x6 = fr["x6"].double().requires_grad_(True)
    be = fr["betas"].double().requires_grad_(True)
    total, jl, pl, _ = oracle.refine_loss(osmpl64, J.double(), sd64, x6, be, fr["gt_mm"].double())
    total.backward()
    ref = jrr.PoseRefiner(smpl_tc, J, critic_sd, use_graph=False)
    st = ref._buffers(40)
    st["x6"].copy_(fr["x6"]); st["betas"].copy_(fr["betas"]); st["gt"].copy_(fr["gt_mm"])
    ref._run_chunk(st, 1, 40)
    torch.cuda.synchronize()
    loss = st["loss"].cpu().double()
    assert abs(loss[1] - jl.item()) / jl.item() < 1e-5
    assert abs(loss[2] - pl.item()) / pl.item() < 1e-5
    assert abs(loss[0] - total.item()) / total.item() < 1e-5
    # Adam's first step: m_hat = g, v_hat = g^2  ->  delta = -lr * g / (|g| + eps)
    for name, g in (("x6", x6.grad), ("betas", be.grad)):
        exp = -1e-2 * g / (g.abs() + 1e-8)
        got = (st[name].cpu().double() - fr[name].double())
        big = g.abs() > 1e-6          # where the direction is well defined
        assert (got[big] - exp[big]).abs().max().item() < 1e-5, name
        # recover the gradient itself from m (first moment = 0.1*g after one step)
    m = st["m"].cpu().double() * 10
    gall = torch.cat([x6.grad.reshape(40, 144), be.grad], dim=1)
    err = (m - gall).abs().max().item() / gall.abs().max().item()
    print(f"[{which}] gradient rel err vs fp64 oracle: {err:.2e}")
    assert err < 1e-4


@pytest.mark.parametrize("use_graph", [False, True])
def test_refine_100_iterations_mpjpe(use_graph, smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped, frames64):
    """64 frames x 100 Adam iterations: refined MPJPE within 0.01 mm of the oracle's."""
    fr = frames64
    x6o, bo, hist = _oracle_refine(oracle, osmpl32, J_shipped, critic_sd, fr, 100)
    Ro = oracle.rot6d_to_rotmat(x6o.reshape(-1, 6)).view(-1, 24, 3, 3)
    mp_o, pa_o = oracle.evaluate(oracle.find_joints(osmpl32, bo, Ro[:, :1], Ro[:, 1:], J_shipped), fr["gt_mm"])
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=use_graph)
    x6 = fr["x6"].to(DEV).clone()
    be = fr["betas"].to(DEV).clone()
    loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=100)
    Rg = jrr.rot6d_to_rotmat(x6.reshape(-1, 6)).view(-1, 24, 3, 3)
    with torch.no_grad():
        pred = jrr.find_joints(smpl_tc, be, Rg[:, :1], Rg[:, 1:], J_shipped.to(DEV))
    mp_g, pa_g = jrr.evaluate(pred, fr["gt_mm"].to(DEV))
    R0 = oracle.rot6d_to_rotmat(fr["x6"].reshape(-1, 6)).view(-1, 24, 3, 3)
    mp_0, _ = oracle.evaluate(oracle.find_joints(osmpl32, fr["betas"], R0[:, :1], R0[:, 1:], J_shipped), fr["gt_mm"])
    print(f"MPJPE initial {mp_0:.3f} mm -> oracle {mp_o:.4f} / cuda {mp_g:.4f} mm; PA {pa_o:.4f} / {pa_g:.4f}; "
          f"final loss oracle {hist[-1][0]:.5f} cuda {loss[0].item():.5f}; "
          f"max |dx6| {(x6.cpu() - x6o).abs().max().item():.2e}")
    assert mp_g < mp_0
    assert abs(mp_g - mp_o) < 0.01
    assert abs(pa_g - pa_o) < 0.01


def test_refine_shard_equals_full_batch(smpl_tc, jrr, critic_sd, J_shipped, frames64):
    """Two shards with logical_batch = full batch reproduce the full-batch trajectory
    bit for bit (frames are independent; only the mean divisors couple them)."""
    fr = frames64
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    gt = fr["gt_mm"].to(DEV)
    xa, ba = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    ref.refine(xa, ba, gt, iters=5)
    xb, bb = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    for lo, hi in ((0, 24), (24, 64)):
        ref.refine(xb[lo:hi], bb[lo:hi], gt[lo:hi], iters=5, logical_batch=64)
    assert torch.equal(xa, xb) and torch.equal(ba, bb)


# ------------------------------------------------------------------ regressor refit
@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_regressor_refit_matches_oracle(which, smpl_tc, jrr, oracle, osmpl32, osmpl64, J_shipped, J_dense, frames64):
    J = J_shipped if which == "shipped" else J_dense
    fr = frames64
    refit = jrr.RegressorRefit(smpl_tc, J, lr=1e-2, chunk=48)      # ragged chunks: 48 + 16
    opt = oracle.RegressorAdam(J.double(), lr=1e-2)
    x6, be, gt = fr["x6"].to(DEV), fr["betas"].to(DEV), fr["gt_mm"].to(DEV)
    for it in range(3):
        g64, l64 = oracle.regressor_grad(osmpl64, opt.J.detach(), fr["x6"].double(), fr["betas"].double(),
                                         fr["gt_mm"].double())
        Jo = opt.step(g64)
        loss = refit.step(x6, be, gt)
        torch.cuda.synchronize()
        assert abs(loss.item() - l64) / l64 < 1e-4
        d = (refit.J_regressor.cpu().double() - Jo).abs().max().item()
        print(f"[{which}] refit step {it}: loss {loss.item():.6e} (oracle {l64:.6e}); max |dJ| {d:.2e}")
        assert d < 1e-4
    # zero entries of the raw regressor never move (relu'(<=0) = 0)
    assert torch.equal(refit.J_regressor.cpu()[J <= 0], J[J <= 0])


def test_artefact_round_trip(tmp_path, jrr, J_shipped):
    """A tensor saved the way the reference artefact was (cuda-tagged, requires_grad,
    column-major) loads unchanged."""
    t = J_shipped.to(DEV).t().contiguous().t().requires_grad_(True)
    assert t.stride() == (1, 17)
    p = tmp_path / "retrained_J_Regressor.pt"
    torch.save(t, p)
    back = jrr.load_j_regressor(str(p), DEV)
    assert back.is_contiguous() and not back.requires_grad and torch.equal(back.cpu(), J_shipped)


# ------------------------------------------------------------------ edge cases, errors, determinism
@pytest.mark.parametrize("B", [1, 127, 129, 300])
def test_ragged_batches_through_the_fused_step(B, smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped):
    """Batches that do not fill the 128-pose tile: 2 fused steps agree with the oracle."""
    fr = make_frames(jrr, oracle, osmpl32, J_shipped, B, 40 + B)
    x6o, bo, hist = oracle.refine(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], fr["gt_mm"], iters=2)
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    x6, be = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=2)
    assert (x6.cpu() - x6o).abs().max().item() < 2e-4
    assert (be.cpu() - bo).abs().max().item() < 2e-4
    assert abs(loss[0].item() - hist[-1][0]) / hist[-1][0] < 1e-4


def test_refinement_is_bitwise_deterministic(smpl_tc, jrr, critic_sd, J_dense, frames64):
    """Fixed-order reductions, no float atomics: two runs are identical bit for bit."""
    outs = []
    for _ in range(2):
        ref = jrr.PoseRefiner(smpl_tc, J_dense, critic_sd, use_graph=True)
        x6, be = frames64["x6"].to(DEV).clone(), frames64["betas"].to(DEV).clone()
        ref.refine(x6, be, frames64["gt_mm"].to(DEV), iters=10)
        outs.append((x6.clone(), be.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_tensor_core_and_simt_paths_agree(smpl_tc, smpl_simt, jrr, critic_sd, J_dense, frames64):
    """Product path (tcgen05 3xTF32, fused epilogue) vs the fp32 SIMT validation kernels."""
    res, losses = [], []
    for smpl in (smpl_tc, smpl_simt):
        ref = jrr.PoseRefiner(smpl, J_dense, critic_sd, use_graph=False)
        x6, be = frames64["x6"].to(DEV).clone(), frames64["betas"].to(DEV).clone()
        loss = ref.refine(x6, be, frames64["gt_mm"].to(DEV), iters=5)
        res.append(x6)
        losses.append(loss.clone())
    d = (res[0] - res[1]).abs()
    print(f"tc vs simt after 5 Adam steps: max |dx6| {d.max().item():.2e}, mean {d.mean().item():.2e}, "
          f"loss {losses[0][0].item():.6f} vs {losses[1][0].item():.6f}")
    # Adam's m/sqrt(v) amplifies rounding differences on parameters whose gradient is ~0, so the
    # bulk statistic and the loss are compared, not the worst element
    assert d.mean().item() < 1e-5
    assert abs(losses[0][0].item() - losses[1][0].item()) / losses[1][0].item() < 1e-4


def test_full_size_properties_4096(smpl_tc, jrr, model, J_dense):
    """Size-independent properties at the benchmark batch (no oracle at this size): regressed
    joints are convex combinations (rows of J-hat sum to 1), so a rigid global rotation of the
    body rotates the pelvis-centred joints, and the fused path equals the module path."""
    B = 4096
    inp = jrr.synthetic.make_pose_inputs(B, 77)
    R = torch.from_numpy(inp["true_rotmat"]).to(DEV)
    b = torch.from_numpy(inp["true_betas"]).to(DEV)
    J = J_dense.to(DEV)
    with torch.no_grad():
        fused = jrr.find_joints(smpl_tc, b, R[:, :1], R[:, 1:], J)
        full, verts = jrr.find_joints(smpl_tc, b, R[:, :1], R[:, 1:], J, return_verts=True)
    assert rel(fused, full) < 1e-5
    c, s = np.cos(1.1), np.sin(1.1)
    Q = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], device=DEV, dtype=torch.float32)
    R2 = R.clone()
    R2[:, 0] = Q @ R[:, 0]
    with torch.no_grad():
        rot = jrr.find_joints(smpl_tc, b, R2[:, :1], R2[:, 1:], J)
    a = jrr.move_pelvis(fused) @ Q.t()
    # the root joint J0(beta) is the centre of rotation; pelvis-centring removes it only if the
    # regressed pelvis is rotated about the same point, which holds for convex combinations
    d = jrr.move_pelvis(rot) - a
    assert d.abs().max().item() < 5e-5
    assert verts.shape == (B, 6890, 3) and torch.isfinite(verts).all()


def test_errors_are_loud(smpl_tc, jrr, critic_sd, J_shipped):
    nat = smpl_tc.native()
    with pytest.raises(jrr.JrrError):
        nat.smpl_forward(torch.zeros(0, 10, device=DEV), torch.zeros(0, 24, 9, device=DEV), 0)      # empty batch
    with pytest.raises(jrr.JrrError):
        nat.find_joints(torch.zeros(2, 10), torch.zeros(2, 24, 9), 0)                               # CPU tensors
    with pytest.raises(jrr.JrrError):
        nat.set_regressor(torch.zeros(17, 100, device=DEV))                                          # wrong shape
    m2 = jrr.SMPL(model_dict=dict(smpl_tc._model_np), create_transl=False).to(DEV)
    fresh = m2.native()
    x6 = torch.zeros(4, 24, 6, device=DEV); be = torch.zeros(4, 10, device=DEV); gt = torch.zeros(4, 17, 3, device=DEV)
    mm = torch.zeros(4, 154, device=DEV); t = torch.zeros(1, dtype=torch.int32, device=DEV)
    with pytest.raises(jrr.JrrError, match="jrr_set_regressor"):
        fresh.refine_step(x6, be, gt, mm, mm.clone(), t, 1e-2, 1e4, 0.0)
    fresh.set_regressor(J_shipped.to(DEV))
    with pytest.raises(jrr.JrrError, match="jrr_critic_load"):
        fresh.refine_step(x6, be, gt, mm, mm.clone(), t, 1e-2, 1e4, 10.0)


# ------------------------------------------------------------------ models that are not 4-sparse
@pytest.fixture(scope="module")
def model6(jrr):
    """Every vertex skinned to SIX joints: two skinning passes (4 + 2 weights) through the same kernels."""
    m = jrr.synthetic.make_smpl_model(0, skin_weights_per_vertex=6)
    assert ((m["lbs_weights"] != 0).sum(1) == 6).all()
    return m


@pytest.fixture(scope="module")
def smpl6(jrr, model6):
    return jrr.SMPL(model_dict=model6, create_transl=False).to(DEV)


def test_six_weight_model_forward_and_backward(smpl6, model6, jrr, oracle):
    """SMPL.forward / backward of a 6-weights-per-vertex model against the fp64 oracle (dense skinning weights there)."""
    o64 = oracle.OracleSMPL(model6, torch.float64)
    B = 70
    inp = jrr.synthetic.make_pose_inputs(B, 31)
    R = torch.from_numpy(inp["true_rotmat"]); betas = torch.from_numpy(inp["true_betas"])
    g = torch.Generator().manual_seed(2)
    wv, wj = torch.randn(B, 6890, 3, generator=g), torch.randn(B, 49, 3, generator=g)

    def run(fn, dt, dev):
        b = betas.to(dev, dt).requires_grad_(True)
        o = R[:, :1].contiguous().to(dev, dt).requires_grad_(True)
        p = R[:, 1:].contiguous().to(dev, dt).requires_grad_(True)
        out = fn(betas=b, body_pose=p, global_orient=o, pose2rot=False)
        ((out.vertices * wv.to(dev, dt)).sum() + (out.joints * wj.to(dev, dt)).sum()).backward()
        return out, b.grad, o.grad, p.grad
    ro, rb, rg, rp = run(o64, torch.float64, "cpu")
    co, cb, cg, cp = run(smpl6, torch.float32, DEV)
    ev, ej = rel(co.vertices, ro.vertices), rel(co.joints, ro.joints)
    eb, eo, ep = rel(cb, rb), rel(cg, rg), rel(cp, rp)
    print(f"6-weight model: vertices {ev:.2e} joints {ej:.2e}; dbetas {eb:.2e} dorient {eo:.2e} dpose {ep:.2e}")
    assert ev < 1e-5 and ej < 1e-5
    assert eb < 1e-4 and eo < 1e-4 and ep < 1e-4
    # and it differs from truncating every vertex to its four largest weights (the check is not vacuous)
    m4 = dict(model6)
    w = model6["lbs_weights"].copy()
    idx = np.argsort(-w, axis=1)[:, 4:]
    np.put_along_axis(w, idx, 0.0, axis=1)
    m4["lbs_weights"] = w
    t = oracle.OracleSMPL(m4, torch.float64)(betas=betas.double(), body_pose=R[:, 1:].double(), global_orient=R[:, :1].double(), pose2rot=False)
    assert rel(t.vertices, ro.vertices) > 1e-3


@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_six_weight_model_refine_and_refit(which, smpl6, model6, jrr, oracle, critic_sd, J_shipped, J_dense):
    """The refinement step (both loss-path formulations) and the regressor refit (both forms) on the 6-weight model:
    find_joints, one-step losses and gradients, and the refit gradient against the fp64 oracle."""
    J = J_shipped if which == "shipped" else J_dense
    o32, o64 = oracle.OracleSMPL(model6), oracle.OracleSMPL(model6, torch.float64)
    n = 150
    fr = make_frames(jrr, oracle, o32, J, n, 77)
    R = fr["true_rotmat"]
    ref_j = oracle.find_joints(o64, fr["true_betas"].double(), R[:, :1].double(), R[:, 1:].double(), J.double())
    with torch.no_grad():
        got_j = jrr.find_joints(smpl6, fr["true_betas"].to(DEV), R[:, :1].to(DEV), R[:, 1:].to(DEV), J.to(DEV))
    assert rel(got_j, ref_j) < 1e-5
    sd64 = {k: v.double() for k, v in critic_sd.items()}
    x6 = fr["x6"].double().requires_grad_(True)
    be = fr["betas"].double().requires_grad_(True)
    total, jl, pl, _ = oracle.refine_loss(o64, J.double(), sd64, x6, be, fr["gt_mm"].double())
    total.backward()
    gall = torch.cat([x6.grad.reshape(n, 144), be.grad], dim=1)
    kink = oracle.critic_kink_frames(critic_sd, fr["x6"])
    gJ, lJ = oracle.regressor_grad(o64, J.double(), fr["x6"].double(), fr["betas"].double(), fr["gt_mm"].double())
    nat = smpl6.native()
    try:
        for path in ("vertex", "folded"):
            ref = jrr.PoseRefiner(smpl6, J, critic_sd, use_graph=False, loss_path=path)
            st = ref._buffers(n)
            st["x6"].copy_(fr["x6"]); st["betas"].copy_(fr["betas"]); st["gt"].copy_(fr["gt_mm"])
            ref._run_chunk(st, 1, n)
            torch.cuda.synchronize()
            loss = st["loss"].cpu().double()
            d = ((st["m"].cpu().double() * 10 - gall).abs().max(1).values / gall.abs().max())[~kink]
            print(f"[{which}/{path}] 6-weight model: loss {loss[0].item():.6f} vs {total.item():.6f}; gradient rel err {d.max().item():.2e}")
            assert abs(loss[0].item() - total.item()) / total.item() < 1e-5 and abs(loss[1].item() - jl.item()) / jl.item() < 1e-5
            assert d.max().item() < 1e-4
            refit = jrr.RegressorRefit(smpl6, J, lr=1e-2, chunk=100)          # ragged chunks: 100 + 50
            opt = oracle.RegressorAdam(J.double(), lr=1e-2)
            Jo = opt.step(gJ)
            l = refit.step(fr["x6"].to(DEV), fr["betas"].to(DEV), fr["gt_mm"].to(DEV))
            dJ = (refit.J_regressor.cpu().double() - Jo).abs().max().item()
            print(f"[{which}/{path}] 6-weight model refit: loss {l.item():.6e} vs {lJ:.6e}; max |dJ| {dJ:.2e}")
            assert abs(l.item() - lJ) / lJ < 1e-4 and dJ < 1e-4
    finally:
        nat.set_loss_path("vertex")


def test_more_than_24_weights_or_simt_build_is_rejected(jrr, model6):
    with pytest.raises(jrr.JrrError, match="tensor-core build"):
        jrr.SMPL(model_dict=model6, create_transl=False, gemm_impl=1).to(DEV).native()


def test_transl_and_default_parameters(smpl_tc, jrr, model, osmpl32):
    """smplx semantics: module parameters are the defaults, transl is added to both outputs."""
    m = jrr.SMPL(model_dict=model, batch_size=2, create_transl=True).to(DEV)
    with torch.no_grad():
        m.transl.copy_(torch.tensor([[0.1, -0.2, 0.3], [1.0, 2.0, 3.0]]))
        m.betas.normal_()
        m.body_pose.normal_(std=0.2)
    out = m()
    ref = osmpl32(betas=m.betas.detach().cpu(), body_pose=m.body_pose.detach().cpu(),
                  global_orient=m.global_orient.detach().cpu(), transl=m.transl.detach().cpu(), pose2rot=True)
    assert rel(out.vertices, ref.vertices) < 1e-5 and rel(out.joints, ref.joints) < 1e-5
    out.vertices.sum().backward()
    assert m.transl.grad is not None and abs(m.transl.grad[0, 0].item() - 6890) < 1e-2


def test_fused_kernels_match_unfused_kernels(jrr, model, critic_sd, J_dense, frames64, monkeypatch):
    """The fused forward (GEMM + skinning + regressor epilogue) and fused backward (generated-
    operand GEMM) against the separate GEMM / skinning kernels: same inputs, 3 Adam steps."""
    res = {}
    for tag, ff, fb in (("fused", "1", "1"), ("unfused", "0", "0")):
        monkeypatch.setenv("JRR_FUSED_FWD", ff)
        monkeypatch.setenv("JRR_FUSED_BWD", fb)
        smpl = jrr.SMPL(model_dict=model, create_transl=False).to(DEV)      # flags are read at model creation
        ref = jrr.PoseRefiner(smpl, J_dense, critic_sd, use_graph=False)
        st = ref._buffers(64)
        st["x6"].copy_(frames64["x6"]); st["betas"].copy_(frames64["betas"]); st["gt"].copy_(frames64["gt_mm"])
        ref._run_chunk(st, 1, 64)
        torch.cuda.synchronize()
        res[tag] = (st["m"].clone() * 10, st["loss"].clone(), ref.launches_per_step)   # m = 0.1 * gradient after one step
    g_f, l_f, n_f = res["fused"]
    g_u, l_u, n_u = res["unfused"]
    err = (g_f - g_u).abs().max().item() / g_u.abs().max().item()
    print(f"fused vs unfused: gradient rel diff {err:.2e}; loss {l_f[0].item():.6f} vs {l_u[0].item():.6f}; "
          f"launches per step {n_f} vs {n_u}")
    assert err < 2e-5
    assert abs(l_f[0].item() - l_u[0].item()) / l_u[0].item() < 1e-6
    assert n_f < n_u


# ------------------------------------------------------------------ widening: 2-D reprojection + camera fit
def _cam_problem(jrr, oracle, osmpl32, J, B, seed):
    fr = make_frames(jrr, oracle, osmpl32, J, B, seed)
    g = torch.Generator().manual_seed(seed)
    cam_true = torch.stack([0.1 * torch.randn(B, generator=g), 0.1 * torch.randn(B, generator=g),
                            40 + 5 * torch.randn(B, generator=g)], dim=1)
    with torch.no_grad():
        R = fr["true_rotmat"]
        joints = oracle.find_joints(osmpl32, fr["true_betas"], R[:, :1], R[:, 1:], J)
        gt2d = oracle.project_2d(joints, cam_true) + 0.5 * torch.randn(B, 17, 2, generator=g)
    cam0 = cam_true + torch.tensor([0.05, -0.05, 3.0]) * torch.randn(B, 3, generator=g)
    return fr, gt2d, cam0


def test_camera_fit_matches_oracle(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped):
    fr, gt2d, cam0 = _cam_problem(jrr, oracle, osmpl32, J_shipped, 48, 5)
    cam_o, loss_o = oracle.camera_fit(osmpl32, J_shipped, fr["x6"], fr["betas"], gt2d, cam0, iters=300)
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    cam = cam0.to(DEV).clone()
    loss = torch.zeros(1, device=DEV)
    ref.native.camera_fit(fr["x6"].to(DEV), fr["betas"].to(DEV), gt2d.to(DEV), cam, 300, 1e-2, loss_out=loss)
    torch.cuda.synchronize()
    d = (cam.cpu() - cam_o).abs().max().item()
    print(f"camera fit (300 Adam steps): max |dcam| {d:.2e}; loss oracle {loss_o:.5f} cuda {loss.item():.5f}")
    assert d < 2e-3 and abs(loss.item() - loss_o) / loss_o < 1e-3


def test_refine_with_2d_term_matches_oracle(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped):
    fr, gt2d, cam0 = _cam_problem(jrr, oracle, osmpl32, J_shipped, 40, 9)
    x6o, bo, co, hist = oracle.refine_2d(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], cam0, fr["gt_mm"], gt2d,
                                         iters=3)
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    x6, be, cam = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone(), cam0.to(DEV).clone()
    loss = ref.refine_2d(x6, be, cam, fr["gt_mm"].to(DEV), gt2d.to(DEV), iters=3)
    torch.cuda.synchronize()
    print(f"refine+2d 3 steps: |dx6| {(x6.cpu() - x6o).abs().max().item():.2e} |dcam| {(cam.cpu() - co).abs().max().item():.2e}; "
          f"loss {loss.cpu().tolist()} vs {hist[-1]}")
    assert (x6.cpu() - x6o).abs().max().item() < 2e-4
    assert (be.cpu() - bo).abs().max().item() < 2e-4
    assert (cam.cpu() - co).abs().max().item() < 2e-4
    for got, exp in zip(loss.cpu().tolist(), hist[-1]):
        assert abs(got - exp) / max(abs(exp), 1e-12) < 1e-4


def test_evaluate_kernel_matches_oracle(jrr, oracle):
    """On-device MPJPE / PA-MPJPE (Jacobi 3x3 SVD per frame) vs the pinned oracle `evaluate`,
    including reflected and noisy configurations."""
    g = torch.Generator().manual_seed(4)
    for B, noise in ((1, 0.05), (300, 0.05), (257, 0.5)):
        pred = 0.3 * torch.randn(B, 17, 3, generator=g)
        tgt = 1000 * (pred + noise * torch.randn(B, 17, 3, generator=g))
        if B > 1:
            tgt[::7, :, 0] *= -1          # mirrored targets exercise the det < 0 branch
        mo, po = oracle.evaluate(pred.double(), tgt.double())
        mg, pg, pf = jrr.evaluate(pred.to(DEV), tgt.to(DEV), per_frame=True)
        assert abs(mg - mo) < 1e-2 * max(1.0, mo * 1e-3) and abs(pg - po) < 1e-2 * max(1.0, po * 1e-3), (B, mg, mo, pg, po)
        assert pf.shape == (B, 2) and abs(pf[:, 0].mean().item() - mo) < 1e-1
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "ref_utils_golden.npz"))
    mg, pg = jrr.evaluate(torch.from_numpy(z["find_joints"]).to(DEV), torch.from_numpy(z["gt_mm"]).to(DEV))
    assert abs(mg - float(z["mpjpe"])) < 1e-3 and abs(pg - float(z["pa_mpjpe"])) < 1e-3


# ------------------------------------------------------------------ row a7: Shape_Discriminator term
def test_shape_critic_forward_matches_golden(smpl_tc, jrr, oracle):
    import numpy as np, os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_shape_critic_golden.npz"))
    S = jrr.Shape_Discriminator().to(DEV)
    S.load_state_dict({k.replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("shape_operations")})
    got = S.bind(smpl_tc.native())(torch.from_numpy(z["betas"]).to(DEV))
    assert got.shape == (6, 1)
    assert np.abs(got.cpu().numpy() - z["shape_scores"]).max() < 1e-6
    b = torch.randn(1000, 10)
    ref = oracle.shape_discriminator_forward(oracle.make_shape_critic_state_dict(0), b)
    assert (S(b.to(DEV)).cpu() - ref).abs().max() < 1e-6
    smpl_tc.native().load_shape_critic(None)


@pytest.mark.parametrize("w_joint,with_pose", [(1.0, False), (10000.0, True)])
def test_refine_step_with_shape_term(w_joint, with_pose, smpl_tc, jrr, oracle, osmpl64, critic_sd, J_shipped):
    """One Adam step with the shape-critic term on: losses and the betas gradient against the fp64
    oracle.  With w_joint = 1 the shape term is a visible share of d loss / d betas."""
    n = 300          # ragged against the 128-frame blocks
    fr = make_frames(jrr, oracle, oracle.OracleSMPL(jrr.synthetic.make_smpl_model(0)), J_shipped, n, 33)
    ssd = oracle.make_shape_critic_state_dict(5)
    sd64 = {k: v.double() for k, v in critic_sd.items()} if with_pose else None
    x6 = fr["x6"].double().requires_grad_(True)
    be = (2 * fr["betas"]).double().requires_grad_(True)
    total, jl, pl, _ = oracle.refine_loss(osmpl64, J_shipped.double(), sd64, x6, be, fr["gt_mm"].double(), w_joint=w_joint,
                                          shape_sd={k: v.double() for k, v in ssd.items()}, w_shape=10.0, logical_batch=n + 20)
    sl = oracle.shape_loss({k: v.double() for k, v in ssd.items()}, be.detach(), n + 20).item()
    total.backward()
    try:
        ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd if with_pose else None, w_joint=w_joint, use_graph=False,
                              shape_critic_state_dict=ssd, w_shape=10.0)
        st = ref._buffers(n)
        st["x6"].copy_(fr["x6"]); st["betas"].copy_(2 * fr["betas"]); st["gt"].copy_(fr["gt_mm"])
        ref._run_chunk(st, 1, n + 20)
        torch.cuda.synchronize()
        loss = st["loss"].cpu().double()
        assert abs(loss[4] - sl) / sl < 1e-5
        assert abs(loss[1] - jl.item()) / jl.item() < 1e-5
        assert abs(loss[0] - total.item()) / total.item() < 1e-5
        m = st["m"].cpu().double() * 10
        gall = torch.cat([x6.grad.reshape(n, 144), be.grad], dim=1)
        err_b = (m[:, 144:] - be.grad).abs().max().item() / be.grad.abs().max().item()
        err = (m - gall).abs().max().item() / gall.abs().max().item()
        # share of the shape term in the betas gradient (so the check is not vacuous)
        be2 = be.detach().clone().requires_grad_(True)
        (10.0 * oracle.shape_loss({k: v.double() for k, v in ssd.items()}, be2, n + 20)).backward()
        share = be2.grad.abs().max().item() / be.grad.abs().max().item()
        print(f"[w_joint={w_joint}] grad rel err {err:.2e}, betas {err_b:.2e}; shape-term share of max |dbeta| {share:.2e}")
        assert err < 1e-4 and err_b < 1e-4
        if w_joint == 1.0:
            assert share > 1e-2
            # without the term the same step gives a different betas gradient
            ref.native.load_shape_critic(None)
            st["x6"].copy_(fr["x6"]); st["betas"].copy_(2 * fr["betas"])
            ref._run_chunk(st, 1, n + 20)
            m0 = st["m"].cpu().double() * 10
            assert (m0[:, 144:] - be.grad).abs().max().item() / be.grad.abs().max().item() > 1e-3
            assert st["loss"].cpu()[4].item() == 0.0
    finally:
        smpl_tc.native().load_shape_critic(None)


def test_refine_with_shape_term_graph_20_iterations(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped, frames64):
    fr = frames64
    ssd = oracle.make_shape_critic_state_dict(5)
    x6o, bo, hist = oracle.refine(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], fr["gt_mm"], iters=20,
                                  shape_sd=ssd, w_shape=10.0)
    try:
        ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, shape_critic_state_dict=ssd, w_shape=10.0)
        x6 = fr["x6"].to(DEV).clone(); be = fr["betas"].to(DEV).clone()
        loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=20).cpu()
        assert abs(loss[0].item() - hist[-1][0]) / hist[-1][0] < 1e-3
        assert abs(loss[4].item() - hist[-1][4]) / hist[-1][4] < 1e-4
        assert (be.cpu() - bo).abs().max() < 2e-3
    finally:
        smpl_tc.native().load_shape_critic(None)


# ------------------------------------------------------------------ row 8f-1: critic training step
def _flat_grad(jrr, g):
    return jrr.native.flatten_critic_state_dict(g)


def _kink_free_frames(sd64, n, gen, scale=0.7):
    """Random rot6d frames none of whose critic pre-activations sits within round-off of a ReLU kink
    (there fp32 may legitimately take the other branch and a whole gradient row changes by that
    frame's term, which says nothing about the kernels)."""
    keep = []
    while sum(k.shape[0] for k in keep) < n:
        x = scale * torch.randn(2 * n, 24, 6, generator=gen)
        xd = x.double()
        p1 = xd @ sd64["conv_operations.0.weight"].reshape(32, 6).t() + sd64["conv_operations.0.bias"]
        p2 = torch.relu(p1) @ sd64["conv_operations.2.weight"].reshape(32, 32).t() + sd64["conv_operations.2.bias"]
        h = torch.relu(p2).reshape(2 * n, 768)
        a1 = h @ sd64["linear_operations.0.weight"].t() + sd64["linear_operations.0.bias"]
        a2 = torch.relu(a1) @ sd64["linear_operations.2.weight"].t() + sd64["linear_operations.2.bias"]
        ok = (a1.abs().min(1).values > 5e-5) & (a2.abs().min(1).values > 5e-5)
        ok &= (p1.abs().reshape(2 * n, -1).min(1).values > 2e-6) & (p2.abs().reshape(2 * n, -1).min(1).values > 2e-6)
        keep.append(x[ok])
    return torch.cat(keep)[:n].contiguous()


@pytest.mark.parametrize("n,lb", [(300, 320), (1000, 1000)])
def test_critic_weight_gradient_matches_fp64_oracle(n, lb, smpl_tc, jrr, oracle, critic_sd):
    """d/dparams of MSE(D(fake),0) + MSE(D(real),1) (optimize.py:276-281) against fp64 autograd, per tensor."""
    g = torch.Generator().manual_seed(n)
    sd64 = {k: v.double() for k, v in critic_sd.items()}
    fake, real = _kink_free_frames(sd64, n, g), _kink_free_frames(sd64, n, g)
    A = oracle.CriticAdam(sd64, oracle.critic_train_loss)
    loss, grads = A.grad(fake.double(), real.double(), lb)
    nat = smpl_tc.native()
    nat.load_critic(critic_sd)
    G = torch.zeros(1840153, device=DEV); L = torch.zeros(1, device=DEV)
    nat.critic_grad_accumulate(fake.to(DEV), 0.0, G, L, logical_batch=lb)
    nat.critic_grad_accumulate(real.to(DEV), 1.0, G, L, logical_batch=lb)
    torch.cuda.synchronize()
    assert abs(L.item() - loss) / loss < 1e-5
    got = jrr.native.unflatten_state_dict(G.cpu(), jrr.native.CRITIC_KEYS, jrr.native.CRITIC_SHAPES)
    worst = 0.0
    for k, ge in grads.items():
        err = (got[k].double() - ge).abs().max().item() / max(ge.abs().max().item(), 1e-12)
        worst = max(worst, err)
        assert err < 1e-4, (k, err)      # 1-element tensors (head biases) are sums with cancellation: fp32 round-off ~4e-5
    print(f"critic weight gradient n={n}: worst per-tensor rel err {worst:.2e}")
    # accumulate semantics / sharding: two shards of `fake` add up to the full call
    G2 = torch.zeros_like(G); G3 = torch.zeros_like(G)
    nat.critic_grad_accumulate(fake.to(DEV), 0.0, G2, None, logical_batch=lb)
    nat.critic_grad_accumulate(fake[:113].to(DEV), 0.0, G3, None, logical_batch=lb)
    nat.critic_grad_accumulate(fake[113:].to(DEV), 0.0, G3, None, logical_batch=lb)
    assert (G2 - G3).abs().max().item() < 1e-5 * G2.abs().max().item()
    # run-to-run identical (no atomics)
    G4 = torch.zeros_like(G)
    nat.critic_grad_accumulate(fake.to(DEV), 0.0, G4, None, logical_batch=lb)
    assert torch.equal(G2, G4)


def test_shape_critic_weight_gradient_matches_fp64_oracle(smpl_tc, jrr, oracle):
    n, lb = 777, 800
    g = torch.Generator().manual_seed(5)
    fake = 1.5 * torch.randn(n, 10, generator=g); real = 1.5 * torch.randn(n, 10, generator=g)
    ssd = oracle.make_shape_critic_state_dict(2)
    A = oracle.CriticAdam({k: v.double() for k, v in ssd.items()}, oracle.shape_critic_train_loss)
    loss, grads = A.grad(fake.double(), real.double(), lb)
    nat = smpl_tc.native()
    try:
        nat.load_shape_critic(ssd, 10.0)
        G = torch.zeros(171, device=DEV); L = torch.zeros(1, device=DEV)
        nat.critic_grad_accumulate(fake.to(DEV), 0.0, G, L, logical_batch=lb, shape=True)
        nat.critic_grad_accumulate(real.to(DEV), 1.0, G, L, logical_batch=lb, shape=True)
        assert abs(L.item() - loss) / loss < 1e-5
        got = jrr.native.unflatten_state_dict(G.cpu(), jrr.native.SHAPE_CRITIC_KEYS, jrr.native.SHAPE_CRITIC_SHAPES)
        for k, ge in grads.items():
            assert (got[k].double() - ge).abs().max().item() / ge.abs().max().item() < 1e-5, k
    finally:
        nat.load_shape_critic(None)


def test_critic_trainer_three_steps_match_oracle_adam(smpl_tc, jrr, oracle, critic_sd):
    """CriticTrainer (optimize.py:113-123,276-293) against torch.optim.Adam on the oracle: losses of
    three consecutive steps (each depends on the previous update), parameters after them, and
    the packed copies used by the refinement kernels are the updated ones."""
    n = 512
    g = torch.Generator().manual_seed(9)
    fake = 0.7 * torch.randn(n, 24, 6, generator=g); real = 0.7 * torch.randn(n, 24, 6, generator=g)
    bf = torch.randn(n, 10, generator=g); br = torch.randn(n, 10, generator=g)
    ssd = oracle.make_shape_critic_state_dict(2)
    A = oracle.CriticAdam(critic_sd, oracle.critic_train_loss, lr=1e-3)
    S = oracle.CriticAdam(ssd, oracle.shape_critic_train_loss, lr=1e-3)
    nat = smpl_tc.native()
    try:
        T = jrr.CriticTrainer(smpl_tc, critic_sd, ssd, lr=1e-3, chunk=200)      # ragged chunks
        for i in range(3):
            lo, ls = A.step(fake, real), S.step(bf, br)
            lp, lsh = T.step(fake.to(DEV), real.to(DEV), bf.to(DEV), br.to(DEV))
            assert abs(lp.item() - lo) / lo < 2e-5, (i, lp.item(), lo)
            assert abs(lsh.item() - ls) / ls < 2e-5, (i, lsh.item(), ls)
        sd_o, sd_c = A.state_dict(), T.state_dict()
        assert list(sd_c.keys()) == list(sd_o.keys())
        for k in sd_o:
            assert sd_c[k].shape == sd_o[k].shape
            d = (sd_c[k].cpu() - sd_o[k]).abs()
            # Adam moves every weight by ~lr per step whatever |g| is, so round-off (and the odd ReLU branch
            # flip) shows where |g| ~ 0: mean within 3 % of the 3*lr travelled, max within 2*3*lr
            assert d.mean().item() < 1e-4 and d.max().item() <= 6.1e-3, (k, d.mean().item(), d.max().item())
        so, sc = S.state_dict(), T.shape_state_dict()
        for k in so:
            assert (sc[k].cpu() - so[k]).abs().max().item() < 1e-5, k
        # the refinement path now scores with the trained weights
        x = 0.7 * torch.randn(50, 24, 6)
        got = nat.critic_forward(x.to(DEV)).cpu()
        ref = oracle.discriminator_forward(sd_c_cpu := {k: v.cpu() for k, v in sd_c.items()}, x)[..., 0]
        assert (got - ref).abs().max().item() < 1e-5
        assert (got - oracle.discriminator_forward(critic_sd, x)[..., 0]).abs().max().item() > 1e-4
    finally:
        nat.load_shape_critic(None)
        nat.load_critic(critic_sd)


# ------------------------------------------------------------------ the whole per-batch loop of optimize.py:150-312
def test_refinement_loop_two_batches_match_oracle_composition(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped):
    """camera fit -> refinement with all in-scope terms -> critic + shape-critic training step -> regressor
    refit, twice (the critics' and the regressor's Adam state persists across batches), against the same
    sequence composed from the oracle's pieces."""
    n, cam_it, ref_it = 48, 40, 4
    ssd = oracle.make_shape_critic_state_dict(5)
    loop = jrr.RefinementLoop(smpl_tc, J_shipped, critic_sd, ssd, refine_iters=ref_it, cam_iters=cam_it)
    A = oracle.CriticAdam(critic_sd, oracle.critic_train_loss, lr=1e-3)
    S = oracle.CriticAdam(ssd, oracle.shape_critic_train_loss, lr=1e-3)
    RA = oracle.RegressorAdam(J_shipped, lr=1e-2)
    J_o = J_shipped.clone()
    try:
        for bi in range(2):
            fr, gt2d, cam0 = _cam_problem(jrr, oracle, osmpl32, J_shipped, n, 40 + bi)
            gt_raw = fr["gt_mm"] + torch.tensor([30.0, -20.0, 10.0])       # not pelvis-centred on input (optimize.py:162)
            out = loop.run_batch({"orient": fr["x6"][:, :1], "pose": fr["x6"][:, 1:], "betas": fr["betas"],
                                  "gt_j3d": gt_raw, "gt_j2d": gt2d, "cam": cam0})
            torch.cuda.synchronize()
            # ---- oracle composition
            gt = oracle.move_pelvis(gt_raw)
            cam_o, _ = oracle.camera_fit(osmpl32, J_o, fr["x6"], fr["betas"], gt2d, cam0, iters=cam_it)
            x6o, bo, co, hist = oracle.refine_2d(osmpl32, J_o, A.state_dict(), fr["x6"], fr["betas"], cam_o, gt, gt2d,
                                                 iters=ref_it, shape_sd=S.state_dict(), w_shape=10.0)
            lc = A.step(x6o, fr["x6"])
            ls = S.step(bo, fr["betas"])
            g, lr_ = oracle.regressor_grad(osmpl32, J_o, x6o, bo, gt)
            J_o = RA.step(g)
            # ---- compare
            assert (out["x6"].cpu() - x6o).abs().max().item() < 5e-4, bi
            assert (out["betas"].cpu() - bo).abs().max().item() < 5e-4, bi
            assert (out["cam"].cpu() - co).abs().max().item() < 5e-3, bi
            assert abs(out["refine_loss"][0].item() - hist[-1][0]) / hist[-1][0] < 1e-3, bi
            assert abs(out["critic_loss"].item() - lc) / lc < 1e-4, (bi, out["critic_loss"].item(), lc)
            assert abs(out["shape_critic_loss"].item() - ls) / ls < 1e-4, (bi, out["shape_critic_loss"].item(), ls)
            assert abs(out["refit_loss"].item() - lr_) / lr_ < 1e-3, (bi, out["refit_loss"].item(), lr_)
            dJ = (loop.refit.J_regressor.cpu() - J_o).abs().max().item()
            print(f"batch {bi}: critic loss {out['critic_loss'].item():.6f} vs {lc:.6f}; refit loss {out['refit_loss'].item():.3e} "
                  f"vs {lr_:.3e}; max |dJ| {dJ:.2e}")
            assert dJ < 1e-4, bi      # north star: regressor weights within 1e-4
        m0, p0 = loop.evaluate(fr["x6"], fr["betas"], gt_raw)
        m1, p1 = loop.evaluate(out["x6"], out["betas"], gt_raw)
        assert m1 < m0
    finally:
        smpl_tc.native().load_shape_critic(None)
        smpl_tc.native().load_critic(critic_sd)
        smpl_tc.native().set_regressor(J_shipped.to(DEV))


# ------------------------------------------------------------------ folded loss path (jrr_set_loss_path)
@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_folded_loss_path_single_step(which, smpl_tc, jrr, oracle, osmpl64, critic_sd, J_shipped, J_dense):
    """One Adam step through the folded operator T = Jhat o skinning o blend: gradients against the fp64
    oracle and against the per-vertex kernels; same losses."""
    J = J_shipped if which == "shipped" else J_dense
    n = 300
    fr = make_frames(jrr, oracle, oracle.OracleSMPL(jrr.synthetic.make_smpl_model(0)), J, n, 61)
    sd64 = {k: v.double() for k, v in critic_sd.items()}
    x6 = fr["x6"].double().requires_grad_(True)
    be = fr["betas"].double().requires_grad_(True)
    total, jl, pl, _ = oracle.refine_loss(osmpl64, J.double(), sd64, x6, be, fr["gt_mm"].double())
    total.backward()
    gall = torch.cat([x6.grad.reshape(n, 144), be.grad], dim=1)
    res = {}
    try:
        for path in ("vertex", "folded"):
            ref = jrr.PoseRefiner(smpl_tc, J, critic_sd, use_graph=False, loss_path=path)
            st = ref._buffers(n)
            st["x6"].copy_(fr["x6"]); st["betas"].copy_(fr["betas"]); st["gt"].copy_(fr["gt_mm"])
            ref._run_chunk(st, 1, n)
            torch.cuda.synchronize()
            res[path] = (st["m"].cpu().double() * 10, st["loss"].cpu().double(), ref.launches_per_step)
    finally:
        smpl_tc.native().set_loss_path("vertex")
    gf, lf, nf = res["folded"]
    gv, lv, nv = res["vertex"]
    err_o = (gf - gall).abs().max().item() / gall.abs().max().item()
    err_v = (gf - gv).abs().max().item() / gv.abs().max().item()
    print(f"[{which}] folded path: gradient rel err vs fp64 oracle {err_o:.2e}, vs per-vertex kernels {err_v:.2e}; "
          f"launches per step {nf} vs {nv}")
    assert err_o < 1e-4 and err_v < 2e-5
    assert abs(lf[1] - jl.item()) / jl.item() < 1e-5 and abs(lf[0] - total.item()) / total.item() < 1e-5
    assert abs(lf[0] - lv[0]) / lv[0] < 1e-6


def test_folded_loss_path_100_iterations_and_all_terms(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped, frames64):
    fr = frames64
    x6o, bo, hist = _oracle_refine(oracle, osmpl32, J_shipped, critic_sd, fr, 100)
    try:
        ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, loss_path="folded")
        x6 = fr["x6"].to(DEV).clone(); be = fr["betas"].to(DEV).clone()
        loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=100)
        pred_o = oracle.find_joints(osmpl32, bo, *(lambda R: (R[:, :1], R[:, 1:]))(oracle.rot6d_to_rotmat(x6o.reshape(-1, 6)).view(-1, 24, 3, 3)), J_shipped)
        m_o, _ = oracle.evaluate(pred_o, fr["gt_mm"])
        R = oracle.rot6d_to_rotmat(x6.cpu().reshape(-1, 6)).view(-1, 24, 3, 3)
        m_c, _ = oracle.evaluate(oracle.find_joints(osmpl32, be.cpu(), R[:, :1], R[:, 1:], J_shipped), fr["gt_mm"])
        print(f"folded path, 100 iterations: MPJPE oracle {m_o:.4f} mm, cuda {m_c:.4f} mm")
        assert abs(m_o - m_c) < 0.01
        assert abs(loss[0].item() - hist[-1][0]) / hist[-1][0] < 1e-3
        # 2-D term + camera group + shape critic through the folded seed kernel
        fr2, gt2d, cam0 = _cam_problem(jrr, oracle, osmpl32, J_shipped, 40, 9)
        ssd = oracle.make_shape_critic_state_dict(5)
        x6o, bo, co, hist = oracle.refine_2d(osmpl32, J_shipped, critic_sd, fr2["x6"], fr2["betas"], cam0, fr2["gt_mm"], gt2d,
                                             iters=3, shape_sd=ssd, w_shape=10.0)
        ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False, shape_critic_state_dict=ssd, loss_path="folded")
        x6, be, cam = fr2["x6"].to(DEV).clone(), fr2["betas"].to(DEV).clone(), cam0.to(DEV).clone()
        loss = ref.refine_2d(x6, be, cam, fr2["gt_mm"].to(DEV), gt2d.to(DEV), iters=3)
        assert (x6.cpu() - x6o).abs().max().item() < 2e-4 and (be.cpu() - bo).abs().max().item() < 2e-4
        assert (cam.cpu() - co).abs().max().item() < 2e-4
        assert abs(loss[0].item() - hist[-1][0]) / hist[-1][0] < 1e-4
    finally:
        smpl_tc.native().load_shape_critic(None)
        smpl_tc.native().set_loss_path("vertex")


def test_folded_loss_path_follows_the_refit(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped):
    """The folded operator is rebuilt inside jrr_regressor_apply / jrr_set_regressor: the whole per-batch loop on
    the folded path against the oracle composition (second batch runs with the refitted regressor)."""
    n, ref_it = 48, 4
    loop = jrr.RefinementLoop(smpl_tc, J_shipped, critic_sd, None, refine_iters=ref_it, loss_path="folded")
    A = oracle.CriticAdam(critic_sd, oracle.critic_train_loss, lr=1e-3)
    RA = oracle.RegressorAdam(J_shipped, lr=1e-2)
    J_o = J_shipped.clone()
    try:
        for bi in range(2):
            fr = make_frames(jrr, oracle, osmpl32, J_shipped, n, 70 + bi)
            out = loop.run_batch({"orient": fr["x6"][:, :1], "pose": fr["x6"][:, 1:], "betas": fr["betas"], "gt_j3d": fr["gt_mm"]})
            x6o, bo, hist = oracle.refine(osmpl32, J_o, A.state_dict(), fr["x6"], fr["betas"], oracle.move_pelvis(fr["gt_mm"]), iters=ref_it)
            A.step(x6o, fr["x6"])
            g, _ = oracle.regressor_grad(osmpl32, J_o, x6o, bo, oracle.move_pelvis(fr["gt_mm"]))
            J_o = RA.step(g)
            assert (out["x6"].cpu() - x6o).abs().max().item() < 5e-4, bi
            assert abs(out["refine_loss"][0].item() - hist[-1][0]) / hist[-1][0] < 1e-3, bi
            assert (loop.refit.J_regressor.cpu() - J_o).abs().max().item() < 1e-4, bi
    finally:
        smpl_tc.native().set_loss_path("vertex")
        smpl_tc.native().load_critic(critic_sd)
        smpl_tc.native().set_regressor(J_shipped.to(DEV))


@pytest.mark.parametrize("path", ["vertex", "folded"])
def test_multi_step_graph_is_bit_identical(path, smpl_tc, jrr, critic_sd, J_shipped, frames64):
    """10 Adam iterations captured in one CUDA graph (device-side step counter) give exactly the bits of
    ten replays of the one-step graph and of eager launches."""
    fr = frames64
    res = []
    try:
        for kw in (dict(use_graph=False), dict(steps_per_graph=1), dict(steps_per_graph=10)):
            ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, loss_path=path, **kw)
            x6 = fr["x6"].to(DEV).clone(); be = fr["betas"].to(DEV).clone()
            loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=23).clone()      # 2 x 10-step graph + 3 single steps
            res.append((x6.clone(), be.clone(), loss))
    finally:
        smpl_tc.native().set_loss_path("vertex")
    for x6, be, loss in res[1:]:
        assert torch.equal(x6, res[0][0]) and torch.equal(be, res[0][1]) and torch.equal(loss, res[0][2])


# ------------------------------------------------------------------ the configuration bench.py headlines
# (C2: 4096 frames, DENSE 17x6890 regressor, loss 10000*joint + 10*critic, both formulations of the loss path) against
# the oracle itself -- the loop of optimize.py:220-265.  Frames couple only through the divisors of the two mean
# losses, so the oracle runs in chunks with logical_batch = 4096 (tests/test_oracle.py::test_shards_reproduce_...).
BENCH_B = 4096


def _frames_chunked(jrr, oracle, osmpl, J, n, seed, chunk=512):
    inp = jrr.synthetic.make_pose_inputs(n, seed)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    t["gt_mm"] = torch.cat([oracle.make_gt(osmpl, J, t["true_rotmat"][lo:lo + chunk], t["true_betas"][lo:lo + chunk],
                                           t["gt_noise"][lo:lo + chunk]) for lo in range(0, n, chunk)])
    return t


@pytest.fixture(scope="module")
def bench_frames(jrr, oracle, osmpl32, J_dense):
    return _frames_chunked(jrr, oracle, osmpl32, J_dense, BENCH_B, 0)


@pytest.fixture(scope="module")
def bench_oracle_fp64_step(oracle, osmpl64, J_dense, critic_sd, bench_frames):
    """fp64 oracle losses and gradient [4096,154] of ONE iteration on the bench inputs."""
    fr, n, chunk = bench_frames, BENCH_B, 512
    sd64 = {k: v.double() for k, v in critic_sd.items()}
    tot = jl = pl = 0.0
    grads = []
    for lo in range(0, n, chunk):
        x6 = fr["x6"][lo:lo + chunk].double().requires_grad_(True)
        be = fr["betas"][lo:lo + chunk].double().requires_grad_(True)
        t, j, p, _ = oracle.refine_loss(osmpl64, J_dense.double(), sd64, x6, be, fr["gt_mm"][lo:lo + chunk].double(),
                                        logical_batch=n)
        t.backward()
        tot, jl, pl = tot + t.item(), jl + j.item(), pl + p.item()
        grads.append(torch.cat([x6.grad.reshape(-1, 144), be.grad], dim=1))
    return tot, jl, pl, torch.cat(grads)


@pytest.mark.parametrize("path", ["folded", "vertex"])
def test_bench_config_one_step_vs_fp64_oracle(path, smpl_tc, jrr, oracle, critic_sd, J_dense, bench_frames, bench_oracle_fp64_step):
    """B = 4096, dense regressor, one Adam iteration from zero state: the three losses within 1e-5 relative and the
    gradient (Adam's first moment after one step is 0.1*g) within 1e-4 of its max against the fp64 oracle."""
    fr = bench_frames
    tot, jl, pl, gall = bench_oracle_fp64_step
    try:
        ref = jrr.PoseRefiner(smpl_tc, J_dense, critic_sd, use_graph=False, loss_path=path, chunk=BENCH_B)
        st = ref._buffers(BENCH_B)
        st["x6"].copy_(fr["x6"]); st["betas"].copy_(fr["betas"]); st["gt"].copy_(fr["gt_mm"])
        ref._run_chunk(st, 1, BENCH_B)
        torch.cuda.synchronize()
    finally:
        smpl_tc.native().set_loss_path("vertex")
    loss = st["loss"].cpu().double()
    m = st["m"].cpu().double() * 10
    # frames with a critic pre-activation within fp32 round-off of a ReLU kink may take the other branch than fp64 does
    # (one unit's whole contribution, ~1e-3 of that frame's critic gradient): tight bound on the rest, loose on those
    kink = oracle.critic_kink_frames(critic_sd, fr["x6"])
    d = (m - gall).abs().max(1).values / gall.abs().max().item()
    err, err_kink = d[~kink].max().item(), (d[kink].max().item() if kink.any() else 0.0)
    print(f"[{path}] B=4096 dense: loss {loss[0].item():.6f} vs fp64 oracle {tot:.6f}; joint {loss[1].item():.4e} vs {jl:.4e}; "
          f"pose {loss[2].item():.6f} vs {pl:.6f}; gradient rel err {err:.2e} ({int(kink.sum())} ReLU-kink frames: {err_kink:.2e})")
    assert abs(loss[0].item() - tot) / tot < 1e-5
    assert abs(loss[1].item() - jl) / jl < 1e-5
    assert abs(loss[2].item() - pl) / pl < 1e-5
    assert err < 1e-4
    assert int(kink.sum()) < BENCH_B // 50 and err_kink < 2e-3


@pytest.mark.parametrize("path", ["folded", "vertex"])
def test_bench_config_10_iterations_vs_fp32_oracle(path, smpl_tc, jrr, oracle, osmpl32, critic_sd, J_dense, bench_frames):
    """B = 4096, dense regressor, 10 Adam iterations through the replayed CUDA graph against the fp32 oracle's
    trajectory (torch.optim.Adam): per-iteration loss, refined parameters."""
    fr, n, chunk = bench_frames, BENCH_B, 512
    xs, bs, hist = [], [], None
    for lo in range(0, n, chunk):
        x6o, bo, h = oracle.refine(osmpl32, J_dense, critic_sd, fr["x6"][lo:lo + chunk], fr["betas"][lo:lo + chunk],
                                   fr["gt_mm"][lo:lo + chunk], iters=10, logical_batch=n)
        xs.append(x6o); bs.append(bo)
        h = torch.tensor(h, dtype=torch.float64)
        hist = h if hist is None else hist + h
    x6o, bo = torch.cat(xs), torch.cat(bs)
    try:
        ref = jrr.PoseRefiner(smpl_tc, J_dense, critic_sd, use_graph=True, loss_path=path, chunk=BENCH_B, steps_per_graph=5)
        x6, be = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
        loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=10).cpu()
    finally:
        smpl_tc.native().set_loss_path("vertex")
    dx, db = (x6.cpu() - x6o).abs(), (be.cpu() - bo).abs()
    moved = (x6o - fr["x6"]).abs().mean().item()
    print(f"[{path}] B=4096 dense, 10 iterations: final loss {loss[0].item():.6f} vs oracle {hist[-1, 0].item():.6f}; "
          f"max |dx6| {dx.max().item():.2e} mean {dx.mean().item():.2e} (mean travel {moved:.2e}); max |dbetas| {db.max().item():.2e}")
    assert abs(loss[0].item() - hist[-1, 0].item()) / hist[-1, 0].item() < 1e-4
    assert abs(loss[1].item() - hist[-1, 1].item()) / hist[-1, 1].item() < 1e-4
    assert dx.mean().item() < 1e-5 and db.mean().item() < 1e-5
    q999 = torch.quantile(dx.flatten()[::7].double(), 0.999).item()
    print(f"[{path}]   99.9th percentile of |dx6| {q999:.2e}")
    # Adam moves every parameter by ~lr per step whatever |g| is: where a gradient is ~0 (the dense regressor makes every
    # joint nearly the vertex centroid, so most rotations are flat directions) round-off decides the SIGN of an lr-sized
    # move.  The worst element is therefore bounded by iterations * lr, not by the arithmetic (measured 2.1e-2 = two
    # steps); the bulk is what shows the kernels: mean 1.1e-6 and the percentile above
    assert q999 < 2e-4
    assert dx.max().item() < 0.3 * 10 * 1e-2 and db.max().item() < 2e-3


def test_dense_regressor_100_iterations_both_paths(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_dense):
    """256 frames x 100 Adam iterations with the DENSE 17x6890 regressor (the reduction bench.py headlines), both
    formulations: refined MPJPE within 0.01 mm of the oracle's, and the recorded bound on the parameter distance."""
    n = 256
    fr = _frames_chunked(jrr, oracle, osmpl32, J_dense, n, 5)
    x6o, bo, hist = oracle.refine(osmpl32, J_dense, critic_sd, fr["x6"], fr["betas"], fr["gt_mm"], iters=100)
    Ro = oracle.rot6d_to_rotmat(x6o.reshape(-1, 6)).view(-1, 24, 3, 3)
    mp_o, pa_o = oracle.evaluate(oracle.find_joints(osmpl32, bo, Ro[:, :1], Ro[:, 1:], J_dense), fr["gt_mm"])
    R0 = oracle.rot6d_to_rotmat(fr["x6"].reshape(-1, 6)).view(-1, 24, 3, 3)
    mp_0, _ = oracle.evaluate(oracle.find_joints(osmpl32, fr["betas"], R0[:, :1], R0[:, 1:], J_dense), fr["gt_mm"])
    try:
        for path in ("folded", "vertex"):
            ref = jrr.PoseRefiner(smpl_tc, J_dense, critic_sd, loss_path=path)
            x6, be = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
            loss = ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=100)
            R = oracle.rot6d_to_rotmat(x6.cpu().reshape(-1, 6)).view(-1, 24, 3, 3)
            mp_c, pa_c = oracle.evaluate(oracle.find_joints(osmpl32, be.cpu(), R[:, :1], R[:, 1:], J_dense), fr["gt_mm"])
            dx = (x6.cpu() - x6o).abs()
            print(f"[{path}] dense, 256 frames x 100 iterations: MPJPE {mp_0:.3f} -> oracle {mp_o:.4f} / cuda {mp_c:.4f} mm, "
                  f"PA {pa_o:.4f} / {pa_c:.4f}; loss {hist[-1][0]:.6f} / {loss[0].item():.6f}; max |dx6| {dx.max().item():.2e} "
                  f"mean {dx.mean().item():.2e}")
            assert mp_c < mp_0
            assert abs(mp_c - mp_o) < 0.01 and abs(pa_c - pa_o) < 0.01
            assert abs(loss[0].item() - hist[-1][0]) / hist[-1][0] < 1e-3
            # flat directions random-walk under Adam (see test_bench_config_10_iterations_vs_fp32_oracle): the function
            # values above are the parity statement; recorded bounds (measured: max 9.5e-2 = ten lr steps, mean 2.9e-4)
            assert dx.max().item() < 0.3 * 100 * 1e-2 and dx.mean().item() < 1e-3
    finally:
        smpl_tc.native().set_loss_path("vertex")


def test_shipped_100_iterations_records_parameter_distance(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped, frames64):
    """The bound test_refine_100_iterations_mpjpe only printed (round 1: 2.9e-4), asserted so a regression shows."""
    fr = frames64
    x6o, bo, _ = oracle.refine(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], fr["gt_mm"], iters=100)
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd)
    x6, be = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    ref.refine(x6, be, fr["gt_mm"].to(DEV), iters=100)
    d = (x6.cpu() - x6o).abs()
    print(f"shipped, 64 frames x 100 iterations: max |dx6| {d.max().item():.2e} mean {d.mean().item():.2e}; "
          f"max |dbetas| {(be.cpu() - bo).abs().max().item():.2e}")
    assert d.max().item() < 1.5e-3 and d.mean().item() < 2e-5
    assert (be.cpu() - bo).abs().max().item() < 1.5e-3


# ------------------------------------------------------------------ advisor findings of round 1
def test_workspace_growth_recaptures_graphs(smpl_tc, jrr, critic_sd, J_shipped, frames64):
    """A captured graph bakes the workspace pointer in; a later call with a larger batch reallocates the shared
    workspace.  The refiner must notice (generation counter) and re-capture instead of replaying into freed memory."""
    fr = frames64
    nat = smpl_tc.native()
    gt = fr["gt_mm"].to(DEV)
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=True)
    xa, ba = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    ref.refine(xa, ba, gt, iters=7)
    gen = nat.ws_generation
    big = jrr.synthetic.make_pose_inputs(1500, 3)
    R = torch.from_numpy(big["true_rotmat"]).to(DEV)
    out = smpl_tc(betas=torch.from_numpy(big["true_betas"]).to(DEV), body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    junk = [torch.full((1 << 22,), float("nan"), device=DEV) for _ in range(8)]      # re-use of the freed block would show
    assert nat.ws_generation > gen
    xb, bb = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    ref.refine(xb, bb, gt, iters=7)
    torch.cuda.synchronize()
    assert torch.equal(xa, xb) and torch.equal(ba, bb)
    assert all(torch.isnan(j).all() for j in junk) and torch.isfinite(out.vertices).all()


def test_find_joints_with_another_regressor_leaves_the_loop_alone(smpl_tc, jrr, oracle, osmpl32, critic_sd, J_shipped, J_dense, frames64):
    """utils.find_joints is stateless in the reference: calling it with another regressor in between (e.g. to compare
    the initial regressor with the retrained one) must not change what PoseRefiner / RegressorRefit work against."""
    fr = frames64
    gt = fr["gt_mm"].to(DEV)
    R = fr["true_rotmat"].to(DEV)
    b = fr["true_betas"].to(DEV)
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd)
    refit = jrr.RegressorRefit(smpl_tc, J_shipped, lr=1e-2)
    xa, ba = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    ref.refine(xa, ba, gt, iters=5)
    Ja = jrr.RegressorRefit(smpl_tc, J_shipped, lr=1e-2)
    la = Ja.step(xa, ba, gt).item()
    # interleave a foreign regressor
    ref2 = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd)
    xb, bb = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    with torch.no_grad():
        other = jrr.find_joints(smpl_tc, b, R[:, :1], R[:, 1:], J_dense.to(DEV))
    expect = oracle.find_joints(osmpl32, fr["true_betas"], fr["true_rotmat"][:, :1], fr["true_rotmat"][:, 1:], J_dense)
    assert rel(other, expect) < 1e-5
    ref2.refine(xb, bb, gt, iters=5)
    with torch.no_grad():
        jrr.find_joints(smpl_tc, b, R[:, :1], R[:, 1:], J_dense.to(DEV))
    lb = refit.step(xb, bb, gt).item()
    assert torch.equal(xa, xb) and torch.equal(ba, bb)
    assert la == lb and torch.equal(Ja.J_regressor, refit.J_regressor)


def test_changed_scalars_and_shape_critic_invalidate_graphs(smpl_tc, jrr, oracle, critic_sd, J_shipped, frames64):
    """lr / loss weights / the shape-critic switch are baked into captured launches: changing them after a capture
    must take effect (graph key), exactly as on the eager path."""
    fr = frames64
    gt = fr["gt_mm"].to(DEV)

    def run(refiner):
        x, b = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
        refiner.refine(x, b, gt, iters=3)
        return x, b
    g = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=True, steps_per_graph=1)
    e = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, use_graph=False)
    run(g)
    g.lr = e.lr = 3e-3
    g.w_joint = e.w_joint = 5000.0
    xg, bg = run(g); xe, be_ = run(e)
    assert torch.equal(xg, xe) and torch.equal(bg, be_)
    try:
        smpl_tc.native().load_shape_critic(oracle.make_shape_critic_state_dict(5), 10.0)
        xg2, bg2 = run(g); xe2, be2 = run(e)
        assert torch.equal(xg2, xe2) and torch.equal(bg2, be2) and not torch.equal(bg2, bg)
    finally:
        smpl_tc.native().load_shape_critic(None)


def test_transl_reaches_unnormalised_extra_joints(jrr, model, oracle):
    """smpl.py:75-76 regresses the 9 extra joints from the TRANSLATED vertices: with J_regressor_extra rows that do not
    sum to 1 the translation arrives scaled by the row sum (advisor finding, round 1)."""
    md = dict(model)
    ex = model["J_regressor_extra"].copy()
    ex[2] *= 1.7
    ex[5] *= 0.4
    md["J_regressor_extra"] = ex
    m = jrr.SMPL(model_dict=md, batch_size=3, create_transl=True).to(DEV)
    o = oracle.OracleSMPL(md)
    g = torch.Generator().manual_seed(1)
    betas, go, bp = torch.randn(3, 10, generator=g), torch.randn(3, 3, generator=g), 0.2 * torch.randn(3, 69, generator=g)
    tr = torch.randn(3, 3, generator=g)
    out = m(betas=betas.to(DEV), global_orient=go.to(DEV), body_pose=bp.to(DEV), transl=tr.to(DEV))
    ref = o(betas=betas, global_orient=go, body_pose=bp, transl=tr)
    assert rel(out.joints, ref.joints) < 1e-5 and rel(out.vertices, ref.vertices) < 1e-5


@pytest.mark.parametrize("which", ["shipped", "dense"])
def test_regressor_refit_through_the_folded_operator(which, smpl_tc, jrr, oracle, osmpl64, J_shipped, J_dense, frames64):
    """With the folded loss path selected the refit gradient is the adjoint of the fold applied to two small GEMMs
    (G_iv = sum_j w_vj (<dT_ji, P_v> + dc_ji), csrc/jrr_model.cu) instead of a pass over the skinned vertices: same G as
    the per-vertex accumulation and as fp64 autograd, same three Adam steps as the oracle (ragged chunks: 48 + 16)."""
    J = J_shipped if which == "shipped" else J_dense
    fr = frames64
    x6, be, gt = fr["x6"].to(DEV), fr["betas"].to(DEV), fr["gt_mm"].to(DEV)
    nat = smpl_tc.native()
    try:
        G = {}
        for path in ("vertex", "folded"):
            nat.set_loss_path(path)
            refit = jrr.RegressorRefit(smpl_tc, J, lr=1e-2, chunk=48)
            refit.accumulate(x6, be, gt)
            torch.cuda.synchronize()
            G[path] = (refit.G.clone(), refit.loss.clone(), nat.launches)
        # fp64 oracle: dL/dJhat by autograd through find_joints with an already-normalised regressor
        Jn = oracle.normalise_regressor(J.double()).requires_grad_(True)
        R = oracle.rot6d_to_rotmat(fr["x6"].double().reshape(-1, 6)).view(-1, 24, 3, 3)
        verts = osmpl64(global_orient=R[:, :1], body_pose=R[:, 1:], betas=fr["betas"].double(), pose2rot=False).vertices
        pred = torch.einsum('jv,bvk->bjk', Jn, verts)
        loss = ((oracle.move_pelvis(pred) - fr["gt_mm"].double() / 1000) ** 2).sum() / (64 * 51)
        loss.backward()
        act = (oracle.normalise_regressor(J) != 0).any(0)              # columns that can receive gradient
        ref = Jn.grad[:, act]
        ev = (G["vertex"][0].cpu().double()[:, act] - ref).abs().max().item() / ref.abs().max().item()
        ef = (G["folded"][0].cpu().double()[:, act] - ref).abs().max().item() / ref.abs().max().item()
        print(f"[{which}] refit gradient vs fp64 oracle: per-vertex {ev:.2e}, folded {ef:.2e}; loss {G['folded'][1].item():.6e} / "
              f"{G['vertex'][1].item():.6e} / {loss.item():.6e}; launches {G['folded'][2]} vs {G['vertex'][2]}")
        assert ev < 1e-4 and ef < 1e-4
        assert abs(G["folded"][1].item() - loss.item()) / loss.item() < 1e-5
        # three Adam steps on the folded path against the oracle optimiser
        refit = jrr.RegressorRefit(smpl_tc, J, lr=1e-2, chunk=48)
        opt = oracle.RegressorAdam(J.double(), lr=1e-2)
        for it in range(3):
            g64, l64 = oracle.regressor_grad(osmpl64, opt.J.detach(), fr["x6"].double(), fr["betas"].double(), fr["gt_mm"].double())
            Jo = opt.step(g64)
            l = refit.step(x6, be, gt)
            d = (refit.J_regressor.cpu().double() - Jo).abs().max().item()
            assert abs(l.item() - l64) / l64 < 1e-4 and d < 1e-4, (it, l.item(), l64, d)
        assert torch.equal(refit.J_regressor.cpu()[J <= 0], J[J <= 0])
    finally:
        nat.set_loss_path("vertex")


def test_loss_history_through_multi_step_graphs(smpl_tc, jrr, critic_sd, J_shipped, frames64):
    """`refine(loss_history=...)`: every iteration's five loss terms, identical whether the iterations run eagerly, as
    one-step graph replays or inside 10-step graphs (each captured iteration writes its own row), into device or pinned
    host memory; the refined parameters are the same bits as without the read-out."""
    fr = frames64
    gt = fr["gt_mm"].to(DEV)
    out = []
    for kw, pinned in ((dict(use_graph=False), False), (dict(steps_per_graph=1), False), (dict(steps_per_graph=10), True)):
        ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, **kw)
        x6, be = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
        hist = torch.zeros(23, 5).pin_memory() if pinned else torch.zeros(23, 5, device=DEV)
        last = ref.refine(x6, be, gt, iters=23, loss_history=hist).clone()
        torch.cuda.synchronize()
        out.append((x6.clone(), hist.cpu().clone(), last.cpu()))
    for x6, hist, last in out[1:]:
        assert torch.equal(x6, out[0][0]) and torch.equal(hist, out[0][1]) and torch.equal(last, out[0][2])
    h = out[0][1]
    assert torch.equal(h[-1], out[0][2]) and (h[:, 0] > 0).all() and h[-1, 1] < h[0, 1]      # the joint term decreases
    ref = jrr.PoseRefiner(smpl_tc, J_shipped, critic_sd, steps_per_graph=10)
    x6, be = fr["x6"].to(DEV).clone(), fr["betas"].to(DEV).clone()
    ref.refine(x6, be, gt, iters=23)
    assert torch.equal(x6, out[0][0])
