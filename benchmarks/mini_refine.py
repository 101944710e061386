#!/usr/bin/env python
"""Smallest refinement run (eager, folded path, one chunk) -- a workload for compute-sanitizer when a kernel faults."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jrr_b200 as jrr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
torch.manual_seed(0)
sd = jrr.Discriminator().state_dict()
J = torch.from_numpy(jrr.synthetic.make_dense_regressor(0))
inp = jrr.synthetic.make_pose_inputs(n, 3)
x6, be = torch.from_numpy(inp["x6"]).to(dev), torch.from_numpy(inp["betas"]).to(dev)
gt = torch.randn(n, 17, 3, device=dev) * 100
ref = jrr.PoseRefiner(smpl, J, sd, use_graph=False, loss_path="folded")
out = ref.refine(x6.clone(), be.clone(), gt, iters=2)
torch.cuda.synchronize()
print("mini refine done", float(ref.last_loss[0]) if hasattr(ref, "last_loss") else "")
