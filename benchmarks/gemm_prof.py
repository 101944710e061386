#!/usr/bin/env python
"""Diagnostic: role timers of the two-row-block critic GEMM (gemm_ts2_kernel, JRR_GEMM_PROF): where each role waits."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402

dev = torch.device("cuda", 0)
smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
nat = smpl.native()
names = ["tma.wait_empty", "tma.total", "mma.wait_tempty", "mma.wait_full", "mma.wait_ready", "mma.total",
         "prod.wait_full", "prod.wait_afree", "prod.total", "epi.wait_tfull", "epi.total", "mma.fence", "mma.issue", "mma.commit"]
for M, N, K, probe in [(4096, 1024, 1024, 0), (4096, 1024, 768, 0)]:
    os.environ["JRR_GEMM_PROBE"] = str(probe)
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev)
    prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    os.environ["JRR_GEMM_PROF"] = str(prof.data_ptr())
    for _ in range(3):
        prof.zero_()
        nat.debug_gemm(A, B, impl=2)
        torch.cuda.synchronize()
    p = prof.view(148, 16).cpu().double()
    used = p[:, 1] > 0
    tiles = (M // 256) * ((N + (127 if N != 768 else 95)) // (128 if N != 768 else 96))
    stages = (tiles + int(used.sum()) - 1) // int(used.sum()) * (K // 32)
    out = {"M": M, "N": N, "K": K, "probe": probe, "ctas": int(used.sum()), "k_blocks_per_cta": stages}
    for i, n in enumerate(names):
        out[n] = round(p[used, i].mean().item())
    out["clk_per_k_block"] = round(out["mma.total"] / stages)
    print(json.dumps(out), flush=True)
os.environ.pop("JRR_GEMM_PROF", None)
os.environ.pop("JRR_GEMM_PROBE", None)

# the four critic GEMMs as the refinement step launches them (their real epilogues)
import ctypes as C
from jrr_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
sd = jrr.Discriminator().state_dict()
J = torch.from_numpy(jrr.synthetic.make_dense_regressor(0))
ref = jrr.PoseRefiner(smpl, J, sd, use_graph=False, loss_path="folded", chunk=4096)
inp = jrr.synthetic.make_pose_inputs(4096, 0)
x6 = torch.from_numpy(inp["x6"]).to(dev); be = torch.from_numpy(inp["betas"]).to(dev); gt = torch.randn(4096, 17, 3, device=dev)
ref.refine(x6.clone(), be.clone(), gt, iters=2)
prof = torch.zeros(4 * 148 * 16, dtype=torch.int64, device=dev)
L.jrr_debug_set_gemm_prof(C.c_void_p(prof.data_ptr()), 4)
ref.refine(x6.clone(), be.clone(), gt, iters=1)
torch.cuda.synchronize()
L.jrr_debug_set_gemm_prof(None, 0)
for k, tag in enumerate(["fwd L1 (K=768, bias+relu)", "fwd L2 (K=1024, bias+relu+head)", "bwd L2 (K=1024, mask)", "bwd L1 (N=768, K=1024)"]):
    p = prof.view(4, 148, 16)[k].cpu().double()
    used = p[:, 1] > 0
    out = {"gemm": tag, "ctas": int(used.sum())}
    for i, n in enumerate(names):
        out[n] = round(p[used, i].mean().item())
    out["epilogue_exposed"] = out["epi.total"] - out["mma.total"]
    print(json.dumps(out), flush=True)
