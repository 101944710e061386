#!/usr/bin/env python
"""Workload for an ncu launch list of ONE refinement iteration with every loss term (3-D joints, pose critic, 2-D
reprojection, silhouette) on 1024 frames at the reference's 224 x 224 silhouettes: module forward -> rasteriser ->
rasteriser backward -> module backward -> fused step (PoseRefiner.refine_silhouette)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402

dev = torch.device("cuda", 0)
n, S = 1024, 224
model = jrr.synthetic.make_smpl_model(0)
smpl = jrr.SMPL(model_dict=model, create_transl=False).to(dev)
inp = jrr.synthetic.make_pose_inputs(n, 5)
faces = jrr.synthetic.make_local_faces(model["v_template"], lbs_weights=model["lbs_weights"])
rend = jrr.Mesh_Renderer(image_size=S, faces=faces)
torch.manual_seed(0)
ref = jrr.PoseRefiner(smpl, torch.rand(17, 6890) + 0.01, jrr.Discriminator().state_dict(), chunk=n, use_graph=False)
x6 = torch.from_numpy(inp["x6"]).to(dev).reshape(n, 24, 6).contiguous()
betas = torch.zeros(n, 10, device=dev)      # (the synthetic shape directions are uncorrelated noise: see bench.run_silhouette)
cam = torch.tensor([0.0, 0.4, 5000.0 / S * 2.3], device=dev).repeat(n, 1).contiguous()
mask = (torch.rand(n, 1, S, S, device=dev) > 0.5).float()
gt, gt2d = torch.zeros(n, 17, 3, device=dev), torch.full((n, 17, 2), 112.0, device=dev)
ref.refine_silhouette(x6, betas, cam, gt, gt2d, mask, rend, iters=2)
torch.cuda.synchronize()
print("done", flush=True)
