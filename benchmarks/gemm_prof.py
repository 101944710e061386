#!/usr/bin/env python
"""Diagnostic: role time stamps of the CTA-pair GEMMs (gemm_pair_kernel) as one refinement step launches them: SM cycles
since kernel entry at which each CTA finished its prologue, issued / received its first operands, issued its last MMA,
saw the accumulator complete, finished the epilogue and the teardown (jrr_debug_set_gemm_prof)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402
from jrr_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
L = _lib.lib()
torch.manual_seed(0)
sd = jrr.Discriminator().state_dict()
J = torch.from_numpy(jrr.synthetic.make_dense_regressor(0))
ref = jrr.PoseRefiner(smpl, J, sd, use_graph=False, loss_path="folded", chunk=4096)
inp = jrr.synthetic.make_pose_inputs(4096, 0)
x6 = torch.from_numpy(inp["x6"]).to(dev)
be = torch.from_numpy(inp["betas"]).to(dev)
gt = torch.randn(4096, 17, 3, device=dev)
ref.refine(x6.clone(), be.clone(), gt, iters=2)
names = ["entry", "prologue_done", "first_stage_issued", "first_operands_landed", "first_k_block_staged", "last_mma_issued",
         "accumulator_complete", "epilogue_issued", "stores_landed", "teardown_done"]
n_launch = 6
prof = torch.zeros(n_launch * 148 * 16, dtype=torch.int64, device=dev)
L.jrr_debug_set_gemm_prof(C.c_void_p(prof.data_ptr()), n_launch)
ref.refine(x6.clone(), be.clone(), gt, iters=1)
torch.cuda.synchronize()
L.jrr_debug_set_gemm_prof(None, 0)
# launch order of one eager step: the critic's four GEMMs (side stream), the two folded GEMMs (main stream) -- by issue order
for k in range(n_launch):
    p = prof.view(n_launch, 148, 16)[k].cpu().double()
    used = p[:, 9] > 0
    if not used.any():
        continue
    out = {"launch": k, "ctas": int(used.sum())}
    for i, n in enumerate(names):
        if i == 0:
            continue
        v = p[used, i]
        out[n] = [round(v.mean().item()), round(v.max().item())]
    print(json.dumps(out), flush=True)
