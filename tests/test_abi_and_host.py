"""CPU-side checks: libjrr.so loads and exports every symbol include/jrr.h declares (no
compute without a GPU), host-side helpers, and the multi-rank refit logic over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol(jrr):
    hdr = open(os.path.join(ROOT, "include", "jrr.h")).read()
    declared = set(re.findall(r"\b(jrr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 16
    L = ctypes.CDLL(jrr.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name
    from jrr_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    assert _lib.lib().jrr_abi_version() == 2


def test_library_is_sm100a_with_tcgen05_and_tma(jrr):
    sass = subprocess.run(["cuobjdump", "-sass", jrr.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_argument_errors_surface_without_a_gpu(jrr):
    from jrr_b200 import _lib
    L = _lib.lib()
    assert L.jrr_model_create(None, None) == 1
    assert b"null" in L.jrr_last_error()
    assert L.jrr_workspace_bytes(None, 10) == 0
    # entry points added in round 2: argument checks come before any CUDA call
    assert L.jrr_set_external_gradient(None, None, None, None) != 0 and b"null model" in L.jrr_last_error()
    assert L.jrr_find_joints_backward(None, 4, None, None, 0, None, None, None, None, 0, None) != 0
    assert L.jrr_silhouette_forward(0, None, 6890, None, None, 13776, 224, 22.3, 1e-4, 1, None, 0, None, None, None, None, 0, None) != 0
    assert b"empty batch" in L.jrr_last_error()
    assert L.jrr_silhouette_forward(2, None, 6890, None, None, 13776, 224, 22.3, 1e-4, 1, None, 0, None, None, None, None, 0, None) != 0
    assert b"null argument" in L.jrr_last_error()
    assert L.jrr_silhouette_forward(2, None, 6890, None, None, 13776, 0, 22.3, 1e-4, 1, None, 0, None, None, None, None, 0, None) != 0
    assert b"image size" in L.jrr_last_error()
    assert L.jrr_silhouette_backward(2, None, 6890, None, None, 13776, None, None, 224, 22.3, -1.0, 1, None, None, None, None, 0, 1.0,
                                     None, None, None, 0, None) != 0
    assert b"sigma" in L.jrr_last_error()
    need = L.jrr_silhouette_workspace_bytes(4, 6890, 13776, 224)
    # ndc + 64-bit z-buffer + per-face corner gradients + per-frame losses (each 256-byte aligned)
    assert need >= 4 * (6890 * 12 + 224 * 224 * 8 + 13776 * 24 + 4) and need < 4 * (6890 * 12 + 224 * 224 * 8 + 13776 * 24 + 4) + 2048


def test_no_cpu_fallback(jrr, model):
    smpl = jrr.SMPL(model_dict=model)
    with pytest.raises(jrr.JrrError):
        smpl(betas=torch.zeros(1, 10), body_pose=torch.zeros(1, 69), global_orient=torch.zeros(1, 3))


def test_smpl_module_interface(jrr, model):
    smpl = jrr.SMPL(model_dict=model, batch_size=2)
    names = dict(smpl.named_buffers())
    for k, shape in (("v_template", (6890, 3)), ("shapedirs", (6890, 3, 10)), ("posedirs", (207, 20670)),
                     ("J_regressor", (24, 6890)), ("lbs_weights", (6890, 24)), ("J_regressor_extra", (9, 6890)),
                     ("parents", (24,))):
        assert tuple(names[k].shape) == shape
    assert smpl.parents[0] == -1 and smpl.joint_map.shape == (49,)
    assert smpl.transl.shape == (2, 3) and smpl.betas.shape == (2, 10)
    import inspect
    params = list(inspect.signature(smpl.forward).parameters)
    assert params[:3] == ["betas", "body_pose", "global_orient"]
    assert jrr.SMPLOutput._fields[:2] == ("vertices", "joints")


def test_shard_range_partitions_frames(jrr):
    for n, w in ((312000, 8), (4096, 3), (7, 8)):
        r = [jrr.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))


def test_flatten_critic_order(jrr, critic_sd):
    flat = jrr.flatten_critic_state_dict(critic_sd)
    assert flat.numel() == 1840153
    assert torch.equal(flat[:192], critic_sd["conv_operations.0.weight"].reshape(-1))
    off = 192 + 32 + 1024 + 32
    assert torch.equal(flat[off:off + 32], critic_sd["linears.0.weight"].reshape(-1))
    assert flat[off + 32] == critic_sd["linears.0.bias"][0]
    assert flat[-1] == critic_sd["linear_operations.4.bias"][0]


def test_synthetic_inputs_are_deterministic(jrr):
    a, b = jrr.synthetic.make_pose_inputs(5, 3), jrr.synthetic.make_pose_inputs(5, 3)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    m1, m2 = jrr.synthetic.make_smpl_model(0), jrr.synthetic.make_smpl_model(0)
    assert all(np.array_equal(m1[k], m2[k]) for k in m1)
    assert list(m1["parents"][:4]) == [-1, 0, 0, 0]


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import jrr_b200 as jrr
from oracle import jrr_oracle as O
from conftest import shipped_regressor, make_frames
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
model = jrr.synthetic.make_smpl_model(0)
osmpl = O.OracleSMPL(model, torch.float64)
J = shipped_regressor().double()
fr = make_frames(jrr, O, O.OracleSMPL(model), shipped_regressor(), 10, 4)
lo, hi = jrr.shard_range(10, rank, world)
g, l = O.regressor_grad(osmpl, J, fr["x6"][lo:hi].double(), fr["betas"][lo:hi].double(),
                        fr["gt_mm"][lo:hi].double(), logical_batch=10)
l = torch.tensor([l], dtype=torch.float64)
dist.all_reduce(g); dist.all_reduce(l)
gf, lf = O.regressor_grad(osmpl, J, fr["x6"].double(), fr["betas"].double(), fr["gt_mm"].double())
assert torch.allclose(g, gf, atol=1e-14), (g - gf).abs().max()
assert abs(l.item() - lf) < 1e-14
opt = O.RegressorAdam(J); Jn = opt.step(g)
gathered = [torch.zeros_like(Jn) for _ in range(world)]
dist.all_gather(gathered, Jn)
assert all(torch.equal(gathered[0], x) for x in gathered)     # replicated optimiser stays in sync
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_refit_allreduce_gloo(tmp_path):
    """world_size-2 gloo run of the refit data flow: shard -> accumulate with the GLOBAL
    divisor -> all-reduce -> identical Adam step on every rank."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2",
               OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


_WORKER_CRITIC = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import jrr_b200 as jrr
from oracle import jrr_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
g = torch.Generator().manual_seed(3)
n = 14
fake = 0.7 * torch.randn(n, 24, 6, generator=g, dtype=torch.float64)
real = 0.7 * torch.randn(n, 24, 6, generator=g, dtype=torch.float64)
sd = {{k: v.double() for k, v in O.make_critic_state_dict(0).items()}}
A = O.CriticAdam(sd, O.critic_train_loss, lr=1e-3)
lo, hi = jrr.shard_range(n, rank, world)
loss, grads = A.grad(fake[lo:hi], real[lo:hi], logical_batch=n)          # this rank's frames, GLOBAL divisor
flat = torch.cat([grads[k].reshape(-1) for k in A.sd])                   # the 7.36 MB flat gradient of the C ABI
l = torch.tensor([loss], dtype=torch.float64)
dist.all_reduce(flat); dist.all_reduce(l)
lf, gf = A.grad(fake, real)
ff = torch.cat([gf[k].reshape(-1) for k in A.sd])
assert flat.numel() == 1840153
assert torch.allclose(flat, ff, atol=1e-13), (flat - ff).abs().max()
assert abs(l.item() - lf) < 1e-13
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_critic_gradient_allreduce_gloo(tmp_path):
    """world_size-2 gloo run of the critic training step's data flow (optimize.py:276-284 sharded): per-rank
    gradient of MSE(D(fake),0)+MSE(D(real),1) with the GLOBAL divisor -> all-reduce == full-batch gradient."""
    script = tmp_path / "worker_critic.py"
    script.write_text(_WORKER_CRITIC.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29518", WORLD_SIZE="2", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


def test_precomputed_split_round_trip(tmp_path, jrr):
    """The reference's on-disk split layout (scripts/data.py:49-69): write, load, index, batch."""
    import pytest
    n = 37
    g = torch.Generator().manual_seed(0)
    lo = 100 + 300 * torch.rand(n, 2, generator=g)
    hi = lo + 200 + 300 * torch.rand(n, 2, generator=g)
    frames = {
        "bboxes": torch.stack([lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1]], dim=1), "betas": torch.randn(n, 10, generator=g),
        "estimated_translation": torch.randn(n, 3, generator=g), "gt_j2d": 1000 * torch.rand(n, 17, 2, generator=g),
        "gt_j3d": 500 * torch.randn(n, 17, 3, generator=g), "intrinsics": torch.eye(3).repeat(n, 1, 1),
        "orient": torch.randn(n, 1, 6, generator=g), "pose": torch.randn(n, 23, 6, generator=g),
    }
    root = str(tmp_path)
    jrr.write_precomputed(root + "/precomputed_val", frames)
    ds = jrr.data_set("validation", root=root)
    assert len(ds) == n and len(ds.images) == n
    item = ds[5]
    assert set(item) == {"bboxes", "betas", "cam", "gt_j2d", "gt_j3d", "intrinsics", "orient", "pose", "inc_gt"}
    assert torch.equal(item["pose"], frames["pose"][5]) and torch.equal(item["cam"], frames["estimated_translation"][5])
    assert item["gt_j2d"].shape == (17, 2) and item["intrinsics"].shape == (3, 3) and bool(item["inc_gt"])
    seen = 0
    for b in ds.batches(16):
        assert b["pose"].shape[1:] == (23, 6) and b["orient"].shape[1:] == (1, 6)
        seen += b["pose"].shape[0]
    assert seen == n
    assert sum(b["pose"].shape[0] for b in ds.batches(16, shuffle=True, seed=1, drop_last=True)) == 32
    assert torch.equal(ds.batch(slice(3, 9))["gt_j3d"], frames["gt_j3d"][3:9])
    with pytest.raises(FileNotFoundError):
        jrr.data_set("train", root=root)
    # Mask R-CNN silhouettes (data.py:113-131): path derived from the frame path, /255, `valid` read before the top-left
    # 2x2 pixels are cleared
    import numpy as np
    names = [f"{root}/S9/imageSequence/cam0/frame_{i:06d}.npy" for i in range(n)]
    jrr.write_precomputed(root + "/precomputed_val", frames, images=names)
    os.makedirs(f"{root}/S9/maskSequence/cam0", exist_ok=True)
    rng = np.random.default_rng(0)
    raw = (rng.random((n, 8, 8)) > 0.5).astype(np.uint8) * 255
    for i in range(n):
        np.save(f"{root}/S9/maskSequence/cam0/frame_{i:06d}.npy", raw[i])
    ds = jrr.data_set("validation", root=root, mask_loader=np.load)
    mask, valid = ds.masks([3, 0, 36])
    assert mask.shape == (3, 1, 8, 8) and mask.max().item() == 1.0
    assert valid.tolist() == [bool(raw[i, 0, 0]) for i in (3, 0, 36)]
    assert mask[:, :, :2, :2].abs().max().item() == 0
    expect = torch.from_numpy(raw[[3, 0, 36]]).float() / 255
    expect[:, :2, :2] = 0
    assert torch.equal(mask[:, 0], expect)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the CUDA arm): stdout is exactly one JSON
    line with the contract's keys, timed on the oracle port, no device traffic."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--frames", "512"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "refine_step_poses_per_sec" and d["unit"] == "poses/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["frames_per_gpu"] == 512 and d["config"]["regressor"] == "dense"      # same keys as the CUDA arm's config
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_export_normalised_regressor_round_trip(tmp_path, jrr, oracle, J_shipped):
    """the pre-normalised form VIBE / MEVA take (test.py:206-208: ReLU, divide by the row sums) --
    values against the pinned oracle normalisation, file round trip in the artefact's torch.save format, and the
    consumer's own normalisation applied on top is the identity (so handing either form to test.py is safe)."""
    import torch
    p = tmp_path / "J_regressor_normalised.pt"
    Jn = jrr.export_normalised_regressor(J_shipped, str(p))
    assert Jn.shape == (17, 6890) and Jn.is_contiguous() and Jn.dtype == torch.float32
    assert torch.equal(Jn, oracle.normalise_regressor(J_shipped, oracle.find_j_reg_mask(J_shipped)))
    assert torch.allclose(Jn.sum(1), torch.ones(17), atol=1e-6) and (Jn >= 0).all()
    assert (Jn[J_shipped <= 0] == 0).all()
    back = jrr.load_j_regressor(str(p))
    assert torch.equal(back, Jn)
    again = torch.relu(back) / torch.relu(back).sum(1, keepdim=True)          # test.py:206-208 on the exported file
    assert (again - Jn).abs().max() < 1e-7
    # accepts the artefact path directly
    raw = tmp_path / "raw.pt"
    jrr.save_j_regressor(J_shipped, str(raw))
    assert torch.equal(jrr.export_normalised_regressor(str(raw)), Jn)
    bad = J_shipped.clone()
    bad[3] = -bad[3].abs()
    import pytest
    with pytest.raises(ValueError, match="no positive entry"):
        jrr.export_normalised_regressor(bad)


def test_bench_chunking_and_regressor_loading_helpers():
    """bench.py's host helpers: balanced chunks (sizes differ by at most one, all frames covered) and the shipped
    artefact through the product loader (byte copy under tests/golden when /root/reference is absent)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n, c in ((312000, 4096), (39000, 4096), (4096, 4096), (5, 4096), (8193, 4096)):
        ch = bench.balanced_chunks(n, c)
        sizes = [b - a for a, b in ch]
        assert ch[0][0] == 0 and ch[-1][1] == n and all(ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
        assert max(sizes) <= c and max(sizes) - min(sizes) <= 1
    assert bench.balanced_chunks(39000, 4096)[0] == (0, 3900)
    J = bench.load_regressor("shipped")
    assert tuple(J.shape) == (17, 6890) and int((J != 0).sum()) == 107
    assert tuple(bench.load_regressor("dense").shape) == (17, 6890)
