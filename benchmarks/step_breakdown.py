#!/usr/bin/env python
"""Diagnostic: what each branch of the refinement step costs when it runs alone (JRR_DEBUG_SKIP, read once per process):
graph-replayed step time of the C2 workload.  Run once per setting."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402

dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
torch.manual_seed(0)
sd = jrr.Discriminator().state_dict()
J = torch.from_numpy(jrr.synthetic.make_dense_regressor(0))
ref = jrr.PoseRefiner(smpl, J, sd, loss_path="folded", chunk=4096)
inp = jrr.synthetic.make_pose_inputs(4096, 0)
x6 = torch.from_numpy(inp["x6"]).to(dev)
be = torch.from_numpy(inp["betas"]).to(dev)
gt = torch.randn(4096, 17, 3, device=dev) * 100
ref.refine(x6.clone(), be.clone(), gt, iters=20)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    xa, ba = x6.clone(), be.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ref.refine(xa, ba, gt, iters=100)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100)
print(json.dumps({"JRR_DEBUG_SKIP": os.environ.get("JRR_DEBUG_SKIP", "0"), "env": {k: v for k, v in os.environ.items() if k.startswith("JRR_")},
                  "ms_per_step": round(best, 4)}), flush=True)
