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
         "prod.wait_full", "prod.wait_afree", "prod.total", "epi.wait_tfull", "epi.total"]
for M, N, K in [(4096, 1024, 1024), (4096, 1024, 768), (4096, 768, 1024)]:
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
    out = {"M": M, "N": N, "K": K, "ctas": int(used.sum()), "k_blocks_per_cta": stages}
    for i, n in enumerate(names):
        out[n] = round(p[used, i].mean().item())
    out["clk_per_k_block"] = round(out["mma.total"] / stages)
    print(json.dumps(out), flush=True)
os.environ.pop("JRR_GEMM_PROF", None)
