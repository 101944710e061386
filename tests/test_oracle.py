"""The restated smplx part of the oracle (PARITY UNPINNED against the real package, which is
neither vendored nor installable): analytic known-answer tests + fp64 gradcheck."""
import numpy as np
import torch


def test_identity_pose_zero_beta_gives_template(osmpl64):
    I = torch.eye(3, dtype=torch.float64).expand(2, 24, 3, 3)
    out = osmpl64(betas=torch.zeros(2, 10, dtype=torch.float64), body_pose=I[:, 1:], global_orient=I[:, :1],
                  pose2rot=False)
    assert torch.allclose(out.vertices[0], osmpl64.v_template, atol=1e-12)
    j24 = osmpl64.J_regressor @ osmpl64.v_template
    # joint_map entry 8 is 'OP MidHip' -> smpl joint 0
    assert torch.allclose(out.joints[0, 8], j24[0], atol=1e-12)
    assert out.joints.shape == (2, 49, 3)


def test_lbs_weights_partition_of_unity(model):
    assert np.allclose(model["lbs_weights"].sum(1), 1.0, atol=1e-6)
    assert ((model["lbs_weights"] != 0).sum(1) == 4).all()


def test_global_rotation_is_rigid(osmpl64):
    c, s = np.cos(0.4), np.sin(0.4)
    R0 = torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float64)
    R = torch.eye(3, dtype=torch.float64).repeat(1, 24, 1, 1)
    R[:, 0] = R0
    out = osmpl64(betas=torch.zeros(1, 10, dtype=torch.float64), body_pose=R[:, 1:], global_orient=R[:, :1],
                  pose2rot=False)
    J0 = (osmpl64.J_regressor @ osmpl64.v_template)[0]
    assert torch.allclose(out.vertices[0], (osmpl64.v_template - J0) @ R0.t() + J0, atol=1e-12)


def test_rodrigues_matches_matrix_exponential(oracle):
    r = torch.tensor([[0.3, -0.2, 0.5], [1e-4, 0.0, 0.0], [2.0, 1.0, -1.5]], dtype=torch.float64)
    R = oracle.batch_rodrigues(r)
    for i in range(3):
        x, y, z = r[i]
        K = torch.tensor([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=torch.float64)
        assert torch.allclose(R[i], torch.linalg.matrix_exp(K), atol=1e-7)


def test_rot6d_identity_and_orthonormal(oracle):
    I = oracle.rot6d_to_rotmat(torch.tensor([[1.0, 0, 0, 1, 0, 0]]))
    assert torch.allclose(I[0], torch.eye(3))
    R = oracle.rot6d_to_rotmat(torch.randn(50, 6, dtype=torch.float64))
    assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3, dtype=torch.float64).expand(50, 3, 3), atol=1e-12)
    assert torch.allclose(torch.det(R), torch.ones(50, dtype=torch.float64))


def test_regressor_rows_sum_to_one_and_mask_is_all_ones(oracle, J_shipped):
    Jn = oracle.normalise_regressor(J_shipped, oracle.find_j_reg_mask(J_shipped))
    assert torch.allclose(Jn.sum(1), torch.ones(17), atol=1e-6)
    assert (Jn >= 0).all()
    assert oracle.find_j_reg_mask(J_shipped).sum().item() == 17 * 6890      # utils.py:183-186 bug


def test_move_pelvis_and_evaluate_known_answers(oracle):
    x = torch.randn(5, 17, 3)
    assert oracle.move_pelvis(x)[:, 0].abs().max() == 0
    mp, pa = oracle.evaluate(x, 1000 * x)
    assert mp < 1e-3 and pa < 1e-2


def test_zero_weight_critic_scores_half(oracle, critic_sd):
    sd = {k: torch.zeros_like(v) for k, v in critic_sd.items()}
    s = oracle.discriminator_forward(sd, torch.randn(3, 24, 6))
    assert torch.allclose(s, torch.full((3, 25, 1), 0.5))
    assert abs(((s - 1) ** 2).mean().item() - 0.25) < 1e-7


def test_smpl_gradcheck_fp64(osmpl64):
    torch.manual_seed(0)
    betas = torch.randn(1, 10, dtype=torch.float64, requires_grad=True)
    aa = (0.3 * torch.randn(1, 72, dtype=torch.float64)).requires_grad_(True)
    sel = torch.tensor([0, 17, 400, 3000, 6889])

    def f(b, a):
        out = osmpl64(betas=b, global_orient=a[:, :3], body_pose=a[:, 3:], pose2rot=True)
        return out.vertices[:, sel], out.joints
    assert torch.autograd.gradcheck(f, (betas, aa), eps=1e-6, atol=1e-7)


def test_refine_reduces_joint_loss(oracle, osmpl32, J_shipped, critic_sd, jrr):
    import conftest
    fr = conftest.make_frames(jrr, oracle, osmpl32, J_shipped, 8, 5)
    _, _, hist = oracle.refine(osmpl32, J_shipped, critic_sd, fr["x6"], fr["betas"], fr["gt_mm"], iters=15)
    assert hist[-1][1] < 0.5 * hist[0][1]


def test_shards_reproduce_full_batch_gradient(oracle, osmpl64, J_shipped, jrr):
    """sum of shard gradients with logical_batch = global batch == full-batch gradient."""
    import conftest
    fr = conftest.make_frames(jrr, oracle, oracle.OracleSMPL(jrr.synthetic.make_smpl_model(0)), J_shipped, 12, 2)
    d = lambda k: fr[k].double()
    g_full, l_full = oracle.regressor_grad(osmpl64, J_shipped.double(), d("x6"), d("betas"), d("gt_mm"))
    g = torch.zeros_like(g_full)
    l = 0.0
    for lo, hi in ((0, 5), (5, 12)):
        gi, li = oracle.regressor_grad(osmpl64, J_shipped.double(), d("x6")[lo:hi], d("betas")[lo:hi],
                                       d("gt_mm")[lo:hi], logical_batch=12)
        g += gi
        l += li
    assert torch.allclose(g, g_full, atol=1e-14) and abs(l - l_full) < 1e-14


def test_shape_term_shards_and_gradient(oracle, osmpl64, J_shipped, jrr):
    """The Shape_Discriminator term is a per-frame sum / logical batch: two shards add up to the
    full-batch loss and gradient (optimize.py:244,249-250)."""
    sd = {k: v.double() for k, v in oracle.make_shape_critic_state_dict(1).items()}
    b = torch.randn(12, 10, dtype=torch.float64, requires_grad=True)
    full = oracle.shape_loss(sd, b)
    g_full, = torch.autograd.grad(full, b)
    parts, grads = [], []
    for lo, hi in ((0, 5), (5, 12)):
        bs = b[lo:hi].detach().clone().requires_grad_(True)
        l = oracle.shape_loss(sd, bs, logical_batch=12)
        parts.append(l.item()); grads.append(torch.autograd.grad(l, bs)[0])
    assert abs(sum(parts) - full.item()) < 1e-12
    assert (torch.cat(grads) - g_full).abs().max() < 1e-14
    assert g_full.abs().max() > 1e-4      # the default-init network is not dead on N(0,1) betas


def test_folded_operator_is_exact(oracle, osmpl64, J_shipped, jrr):
    """The folded formulation (regressor o skinning o blend operator) is the same function as
    find_joints on the per-vertex body model: fp64 agreement to round-off, for a sparse and a dense regressor."""
    g = torch.Generator().manual_seed(8)
    R = oracle.rot6d_to_rotmat(torch.randn(7 * 24, 6, generator=g).double()).view(7, 24, 3, 3)
    b = torch.randn(7, 10, generator=g).double()
    for J in (J_shipped.double(), torch.from_numpy(jrr.synthetic.make_dense_regressor(0)).double()):
        ref = oracle.find_joints(osmpl64, b, R[:, :1], R[:, 1:], J)
        T, c = oracle.fold_operator(osmpl64, J)
        assert T.shape == (24, 17, 3, 218) and c.shape == (24, 17)
        got = oracle.find_joints_folded(osmpl64, b, R, T, c)
        assert (got - ref).abs().max().item() < 1e-12


def test_independent_lbs_statement_agrees(oracle, osmpl64, model, jrr):
    """oracle.lbs vs the second statement written from the SMPL paper (oracle/lbs_independent.py: NumPy
    per-vertex loops, matrix-exponential rotations, inverted rest-pose transforms; no shared code).  The smplx
    boundary cannot be pinned offline; two independent formulations agreeing is what can be shown instead."""
    from oracle import lbs_independent as I2
    inp = jrr.synthetic.make_pose_inputs(2, 123)
    g = torch.Generator().manual_seed(7)
    aa = torch.cat([torch.randn(2, 1, 3, generator=g, dtype=torch.float64),
                    0.4 * torch.randn(2, 23, 3, generator=g, dtype=torch.float64)], dim=1)
    betas = torch.from_numpy(inp["true_betas"]).double()
    out = osmpl64(betas=betas, body_pose=aa[:, 1:].reshape(2, 69), global_orient=aa[:, 0], pose2rot=True)
    for b in range(2):
        rot = np.stack([I2.rotation_from_axis_angle(aa[b, k].numpy()) for k in range(24)])
        verts, pj = I2.smpl_forward_one(model, betas[b].numpy(), rot)
        j49 = I2.joints49_one(model, verts, pj)
        ev = np.abs(verts - out.vertices[b].numpy()).max() / np.abs(verts).max()
        ej = np.abs(j49 - out.joints[b].numpy()).max() / np.abs(j49).max()
        # smplx's Rodrigues adds 1e-8 inside the norm: the two rotations differ by ~1e-8 relative
        assert ev < 5e-8 and ej < 5e-8, (ev, ej)
    # rotation matrices handed over directly (pose2rot=False, the refinement loop's call): no Rodrigues in between
    R = torch.from_numpy(inp["true_rotmat"]).double()
    out = osmpl64(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    verts, pj = I2.smpl_forward_one(model, betas[0].numpy(), R[0].numpy())
    assert np.abs(verts - out.vertices[0].numpy()).max() < 1e-12
    assert np.abs(I2.joints49_one(model, verts, pj) - out.joints[0].numpy()).max() < 1e-12


def test_critic_kink_frames_flags_exactly_the_frames_near_a_relu_kink(oracle, critic_sd):
    """The helper behind the gradient-parity tests: a frame whose layer-2 pre-activation is moved onto a ReLU kink is
    flagged, its neighbours are not, and on random frames the flagged share stays ~1 %."""
    g = torch.Generator().manual_seed(11)
    x = 0.7 * torch.randn(400, 24, 6, generator=g)
    base = oracle.critic_kink_frames(critic_sd, x)
    assert base.float().mean().item() < 0.05
    # shift the bias of unit 7 of the second wide layer so that frame 3's pre-activation there is 1e-8
    sd = {k: v.clone().double() for k, v in critic_sd.items()}
    xd = x.double()
    h = torch.relu(torch.relu(xd @ sd["conv_operations.0.weight"].reshape(32, 6).t() + sd["conv_operations.0.bias"])
                   @ sd["conv_operations.2.weight"].reshape(32, 32).t() + sd["conv_operations.2.bias"]).reshape(400, 768)
    a1 = torch.relu(h @ sd["linear_operations.0.weight"].t() + sd["linear_operations.0.bias"])
    a2 = a1 @ sd["linear_operations.2.weight"].t() + sd["linear_operations.2.bias"]
    sd["linear_operations.2.bias"][7] -= a2[3, 7] - 1e-8
    flagged = oracle.critic_kink_frames(sd, x)
    assert bool(flagged[3])
    assert int((flagged & ~base).sum()) <= 3      # moving one bias puts (almost) only frame 3 on a kink
