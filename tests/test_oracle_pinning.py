"""Pins the oracle's restatement of scripts/utils.py, eval_utils.py and discriminator.py:
(a) against the committed golden vectors generated from the reference's own functions
(tests/golden/make_golden.py), everywhere; (b) against the reference functions imported by
path, when /root/reference exists (build container)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE_ROOT


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ref_utils_golden.npz"))


def test_golden_rot6d(oracle, gold):
    R = oracle.rot6d_to_rotmat(torch.from_numpy(gold["x6"]).reshape(-1, 6)).view(6, 24, 3, 3)
    assert np.abs(R.numpy() - gold["rotmat"]).max() < 1e-6


def test_golden_find_joints_move_pelvis_evaluate(oracle, osmpl32, J_shipped, gold):
    R = torch.from_numpy(gold["rotmat"])
    pred = oracle.find_joints(osmpl32, torch.from_numpy(gold["betas"]), R[:, :1], R[:, 1:], J_shipped,
                              mask=oracle.find_j_reg_mask(J_shipped))
    assert np.abs(pred.numpy() - gold["find_joints"]).max() < 2e-6
    assert np.abs(oracle.move_pelvis(pred).numpy() - gold["move_pelvis"]).max() < 2e-6
    mp, pa = oracle.evaluate(pred, torch.from_numpy(gold["gt_mm"]))
    assert abs(mp - float(gold["mpjpe"])) < 1e-3 and abs(pa - float(gold["pa_mpjpe"])) < 1e-3
    assert float(gold["mask_sum"]) == 17 * 6890


def test_golden_critic(oracle, critic_sd, gold):
    s = oracle.discriminator_forward(critic_sd, torch.from_numpy(gold["x6"]))
    assert np.abs(s.numpy() - gold["critic_scores"]).max() < 1e-6


def test_golden_shape_critic(oracle):
    """Shape_Discriminator (scripts/discriminator.py:57-74): scores of the reference module, default
    init under seed 0, against the oracle restatement on the reference's own weights AND on the
    oracle's re-creation of that init."""
    z = np.load(os.path.join(GOLDEN, "ref_shape_critic_golden.npz"))
    sd = {k.replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k.startswith("shape_operations")}
    mine = oracle.make_shape_critic_state_dict(0)
    assert list(mine.keys()) == [f"shape_operations.{i}.{n}" for i in (0, 2, 4) for n in ("weight", "bias")]
    assert all(torch.equal(sd[k], mine[k]) for k in mine)
    s = oracle.shape_discriminator_forward(mine, torch.from_numpy(z["betas"]))
    assert s.shape == (6, 1)
    assert np.abs(s.numpy() - z["shape_scores"]).max() < 1e-6


def test_golden_critic_training_steps(oracle, critic_sd):
    """optimize.py:276-293 run with the reference's own modules and torch.optim.Adam (two steps,
    fixture from make_golden.py) against the oracle's CriticAdam."""
    z = np.load(os.path.join(GOLDEN, "ref_critic_train_golden.npz"))
    t = lambda k: torch.from_numpy(z[k])
    A = oracle.CriticAdam(critic_sd, oracle.critic_train_loss, lr=1e-3)
    S = oracle.CriticAdam(oracle.make_shape_critic_state_dict(0), oracle.shape_critic_train_loss, lr=1e-3)
    for i in range(2):
        assert abs(A.step(t("x6_fake"), t("x6_real")) - z["losses"][i]) < 1e-6
        assert abs(S.step(t("betas_fake"), t("betas_real")) - z["losses_shape"][i]) < 1e-6
    sd = A.state_dict()
    for key, name in (("conv0_w", "conv_operations.0.weight"), ("conv2_b", "conv_operations.2.bias"),
                      ("lin3_w", "linears.3.weight"), ("lin3_b", "linears.3.bias"),
                      ("b2", "linear_operations.2.bias"), ("w3", "linear_operations.4.weight")):
        assert np.abs(sd[name].numpy() - z[key]).max() < 2e-6, name
    assert np.abs(sd["linear_operations.0.weight"][:4, :8].numpy() - z["w1_block"]).max() < 2e-6
    for k, v in S.state_dict().items():
        assert np.abs(v.numpy() - z["shape__" + k.replace(".", "__")]).max() < 2e-6, k


def test_golden_regressor_fixture_matches_documented_artefact(J_shipped):
    z = np.load(os.path.join(GOLDEN, "j_regressor_nnz.npz"))
    assert str(z["sha256"]) == "4ea32d1b3b9a135130722218f87eadfcf78321cf2ca6954e14f780eb9b60d079"
    assert int((J_shipped != 0).sum()) == 107 and int((J_shipped > 0).sum()) == 62


needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE_ROOT, "scripts")),
                               reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref(oracle):
    return oracle.load_reference_modules(REFERENCE_ROOT)


@needs_ref
def test_artefact_loads_unchanged(jrr, J_shipped):
    path = os.path.join(REFERENCE_ROOT, "models", "retrained_J_Regressor.pt")
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == \
        "4ea32d1b3b9a135130722218f87eadfcf78321cf2ca6954e14f780eb9b60d079"
    J = jrr.load_j_regressor(path)
    assert J.shape == (17, 6890) and J.is_contiguous() and not J.requires_grad
    assert torch.equal(J, J_shipped)


@needs_ref
def test_live_reference_utils(ref, oracle, osmpl32, J_shipped):
    u, d, e = ref
    x = torch.randn(40, 6)
    assert torch.equal(u.rot6d_to_rotmat(x), oracle.rot6d_to_rotmat(x))
    R = oracle.rot6d_to_rotmat(torch.randn(5 * 24, 6)).view(5, 24, 3, 3)
    b = torch.randn(5, 10)
    a = u.find_joints(osmpl32, b, R[:, :1], R[:, 1:], J_shipped, mask=u.find_j_reg_mask(J_shipped))
    o = oracle.find_joints(osmpl32, b, R[:, :1], R[:, 1:], J_shipped, mask=oracle.find_j_reg_mask(J_shipped))
    assert (a - o).abs().max() < 1e-6
    assert torch.equal(u.move_pelvis(a), oracle.move_pelvis(a))
    gt = 1000 * oracle.move_pelvis(a) + 5 * torch.randn(5, 17, 3)
    (m1, p1), (m2, p2) = u.evaluate(a, gt), oracle.evaluate(a, gt)
    assert abs(m1 - m2) < 1e-3 and abs(p1 - p2) < 1e-3
    S1, S2 = torch.randn(4, 17, 3), torch.randn(4, 17, 3)
    assert (e.batch_compute_similarity_transform_torch(S1, S2) - oracle.procrustes(S1, S2)).abs().max() < 1e-4


@needs_ref
def test_live_reference_discriminator_and_state_dict_layout(ref, oracle, jrr, critic_sd):
    u, d, e = ref
    torch.manual_seed(0)
    D = d.Discriminator()
    sd = D.state_dict()
    assert list(sd.keys()) == list(critic_sd.keys())
    assert all(torch.equal(sd[k], critic_sd[k]) for k in sd)
    x = torch.randn(7, 24, 6)
    assert (D(x) - oracle.discriminator_forward(critic_sd, x)).abs().max() < 1e-6
    # the product-side mirror accepts the reference state_dict unchanged
    mine = jrr.Discriminator()
    mine.load_state_dict(sd)
    assert jrr.flatten_critic_state_dict(mine.state_dict()).numel() == 1840153
    torch.manual_seed(3)
    S = d.Shape_Discriminator()
    b = torch.randn(9, 10)
    assert (S(b) - oracle.shape_discriminator_forward(S.state_dict(), b)).abs().max() < 1e-6
    mine_s = jrr.Shape_Discriminator()
    mine_s.load_state_dict(S.state_dict())
    assert jrr.native.flatten_shape_critic_state_dict(mine_s.state_dict()).numel() == 171


@needs_ref
def test_product_mirrors_match_reference_utils(ref, jrr):
    u, d, e = ref
    x = torch.randn(33, 6)
    assert torch.allclose(u.rot6d_to_rotmat(x), jrr.rot6d_to_rotmat(x), atol=1e-7)
    j = torch.randn(4, 17, 3)
    assert torch.equal(u.move_pelvis(j), jrr.move_pelvis(j))
    gt = 1000 * j + torch.randn(4, 17, 3)
    (m1, p1), (m2, p2) = u.evaluate(j, gt), jrr.evaluate(j, gt)
    assert abs(m1 - m2) < 1e-3 and abs(p1 - p2) < 1e-3
    J = torch.randn(17, 6890)
    assert torch.equal(u.find_j_reg_mask(J), jrr.find_j_reg_mask(J))


def test_golden_data_crop_arithmetic(jrr):
    """scripts/data.py find_crop / crop_intrinsics / resize_intrinsics and the 2-D joint repositioning of
    data_set.__getitem__ (fixture produced by the reference's own find_crop)."""
    z = np.load(os.path.join(GOLDEN, "ref_data_crop_golden.npz"))
    t = lambda k: torch.from_numpy(z[k])
    mnx, mny, sc, _, _ = jrr.data.crop_window(t("bboxes"))
    assert np.abs(mnx.numpy() - z["min_x"]).max() < 1e-3 and np.abs(mny.numpy() - z["min_y"]).max() < 1e-3
    assert np.abs(sc.numpy() - z["scale"]).max() < 1e-6
    assert np.abs(jrr.data.cropped_intrinsics(t("intrinsics"), t("bboxes")).numpy() - z["intrinsics_out"]).max() < 1e-3
    assert np.abs(jrr.data.reposition_j2d(t("gt_j2d"), t("bboxes")).numpy() - z["gt_j2d_repositioned"]).max() < 1e-3


@needs_ref
def test_live_reference_find_crop(oracle, jrr):
    rd = oracle.load_reference_data_module()
    g = torch.Generator().manual_seed(3)
    lo = 50 + 400 * torch.rand(9, 2, generator=g)
    hi = lo + 100 + 300 * torch.rand(9, 2, generator=g)
    bb = torch.stack([lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1]], dim=1)
    intr = torch.eye(3).repeat(9, 1, 1)
    intr[:, 0, 0] = 1100; intr[:, 1, 1] = 1090; intr[:, 0, 2] = 505; intr[:, 1, 2] = 498
    for size in (224, 256):
        _, mnx, mny, sc, io = rd.find_crop(torch.zeros(9, 3, 32, 32), bb, intr, img_size=size)
        a, b, c, _, _ = jrr.data.crop_window(bb)
        assert torch.allclose(a, mnx, atol=1e-3) and torch.allclose(b, mny, atol=1e-3) and torch.allclose(c, sc)
        assert torch.allclose(jrr.data.cropped_intrinsics(intr, bb, img_size=size), io, atol=1e-3)


def test_committed_artefact_bytes_load_unchanged(jrr, J_shipped):
    """tests/golden/retrained_J_Regressor.pt is the reference artefact byte for byte (make_golden.py), so this runs on
    the GPU box too, where /root/reference is absent: cuda:0 device tag, requires_grad and stride (1,17) all go through
    load_j_regressor (test.py:46-47 loads it without map_location and needs a GPU for that)."""
    path = os.path.join(GOLDEN, "retrained_J_Regressor.pt")
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == \
        "4ea32d1b3b9a135130722218f87eadfcf78321cf2ca6954e14f780eb9b60d079"
    J = jrr.load_j_regressor(path)
    assert J.shape == (17, 6890) and J.is_contiguous() and not J.requires_grad and J.dtype == torch.float32
    assert torch.equal(J, J_shipped)
    assert int((J != 0).sum()) == 107 and int((J > 0).sum()) == 62
