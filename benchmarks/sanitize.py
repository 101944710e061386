#!/usr/bin/env python
"""Workload for `compute-sanitizer --tool memcheck python benchmarks/sanitize.py` (small batches, every
entry point of include/jrr.h once, both loss-path formulations, ragged sizes).  Prints "sanitizer workload done"."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jrr_b200 as jrr  # noqa: E402
from conftest import shipped_regressor  # noqa: E402

dev = torch.device("cuda:0")
model = jrr.synthetic.make_smpl_model(0)
smpl = jrr.SMPL(model_dict=model, create_transl=False).to(dev)
torch.manual_seed(0)
sd = jrr.Discriminator().state_dict()
ssd = jrr.Shape_Discriminator().state_dict()
for J in (shipped_regressor(), torch.rand(17, 6890) + 0.01):
    for n in (5, 300):
        inp = jrr.synthetic.make_pose_inputs(n, 3)
        x6, be = torch.from_numpy(inp["x6"]).to(dev), torch.from_numpy(inp["betas"]).to(dev)
        R = torch.from_numpy(inp["true_rotmat"]).to(dev)
        out = smpl(betas=be.clone().requires_grad_(True), body_pose=R[:, 1:].clone().requires_grad_(True),
                   global_orient=R[:, :1], pose2rot=False)
        (out.vertices.sum() + out.joints.sum()).backward()
        gt = 1000 * jrr.move_pelvis(jrr.find_joints(smpl, be, R[:, :1], R[:, 1:], J.to(dev))) + 5 * torch.randn(n, 17, 3, device=dev)
        cam = torch.tensor([0.0, 0.0, 40.0], device=dev).repeat(n, 1)
        gt2d = 112 + 20 * torch.randn(n, 17, 2, device=dev)
        for path in ("vertex", "folded"):
            loop = jrr.RefinementLoop(smpl, J, sd, ssd, refine_iters=2, cam_iters=3, loss_path=path)
            for use2d in (True, False):
                b = {"orient": x6[:, :1], "pose": x6[:, 1:], "betas": be, "gt_j3d": gt}
                if use2d:
                    b.update(gt_j2d=gt2d, cam=cam)
                res = loop.run_batch(b)
            loop.evaluate(res["x6"], res["betas"], gt)
            # graph-captured refinement, 10 iterations per graph
            jrr.PoseRefiner(smpl, J, sd, loss_path=path).refine(x6.clone(), be.clone(), gt, iters=12)
        smpl.native().set_loss_path("vertex")
        jrr.Discriminator().to(dev).bind(smpl.native())(x6)
        jrr.Shape_Discriminator().to(dev).bind(smpl.native())(be)
        smpl.native().load_shape_critic(None)
# single-launch small-batch forward (all three rotation formats) and the silhouette term
for n in (1, 7):
    inp = jrr.synthetic.make_pose_inputs(n, 4)
    x6, be = torch.from_numpy(inp["x6"]).to(dev), torch.from_numpy(inp["betas"]).to(dev)
    R = torch.from_numpy(inp["true_rotmat"]).to(dev)
    smpl(betas=be, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    smpl(betas=be, body_pose=torch.randn(n, 69, device=dev), global_orient=torch.randn(n, 3, device=dev))
    smpl.native().smpl_forward(be, x6.reshape(n, 24, 6).contiguous(), 2, False, True)
    rend = jrr.Mesh_Renderer(image_size=40, faces=jrr.synthetic.make_local_faces(model["v_template"], lbs_weights=model["lbs_weights"]))
    cam = torch.tensor([0.0, 0.4, 5000.0 / 40 * 2.3], device=dev).repeat(n, 1)
    bq = be.clone().requires_grad_(True)
    img = jrr.render_mesh(smpl, rend, bq, R[:, :1], R[:, 1:], {"cam": cam})
    torch.nn.functional.mse_loss(img, torch.rand_like(img)).backward()
    gt = torch.zeros(n, 17, 3, device=dev)
    jrr.PoseRefiner(smpl, shipped_regressor(), sd, use_graph=False).refine_silhouette(
        x6.clone(), be.clone(), cam.clone(), gt, torch.full((n, 17, 2), 112.0, device=dev), torch.rand(n, 1, 40, 40, device=dev), rend, iters=2)
torch.cuda.synchronize()
print("sanitizer workload done")
