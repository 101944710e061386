#!/usr/bin/env python
"""Workload for an ncu capture of the small-batch module path: SMPL forward and backward at B = 8 and 1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402

dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
nat = smpl.native()
base = jrr.synthetic.make_pose_inputs(8, 7)
for B in (8, 1):
    full = torch.from_numpy(base["true_rotmat"])[:B].reshape(B, 24, 9).to(dev).contiguous()
    b = torch.from_numpy(base["true_betas"])[:B].to(dev).contiguous()
    dv, dj = torch.randn(B, 6890, 3, device=dev), torch.randn(B, 49, 3, device=dev)
    for _ in range(2):
        nat.smpl_forward(b, full, 0, True, True)
        nat.smpl_backward(b, full, 0, dv, dj)
torch.cuda.synchronize()
print("done", flush=True)
