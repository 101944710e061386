#!/usr/bin/env python
"""Diagnostic: the CTA-pair critic GEMM (gemm_pair_kernel, JRR_GEMM_PAIR) against fp64 and against the other
A-through-TMEM kernels, per critic shape.  The dispatch mode is read once per process: run it once per JRR_GEMM_PAIR value.
Kernel time = (t[11 launches] - t[1 launch]) / 10, so the operand split kernel of the debug entry point cancels."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    smpl = jrr.SMPL(model_dict=jrr.synthetic.make_smpl_model(0), create_transl=False).to(dev)
    nat = smpl.native()
    mode = os.environ.get("JRR_GEMM_PAIR", "1")
    shapes = [(256, 256, 224), (384, 1024, 768), (512, 768, 1024), (4096, 1024, 768), (4096, 1024, 1024), (4096, 768, 1024),
              (8192, 1024, 1024), (9472, 1024, 1024)]
    for M, N, K in shapes:
        g = torch.Generator(device="cpu").manual_seed(M + N + K)
        A = torch.randn(M, K, generator=g).to(dev)
        B = torch.randn(N, K, generator=g).to(dev)
        os.environ["JRR_GEMM_PROBE_REPS"] = "1"
        C = nat.debug_gemm(A, B, impl=2)
        torch.cuda.synchronize()
        ref = A.double() @ B.double().t()
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        same = bool(torch.equal(C, nat.debug_gemm(A, B, impl=2)))
        ts = {}
        for reps in (1, 11):
            os.environ["JRR_GEMM_PROBE_REPS"] = str(reps)
            for _ in range(3):
                nat.debug_gemm(A, B, impl=2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(7):
                e0.record()
                nat.debug_gemm(A, B, impl=2)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ts[reps] = best
        us = (ts[11] - ts[1]) / 10 * 1e3
        print(json.dumps({"pair_mode": mode, "M": M, "N": N, "K": K, "rel_err": err, "bit_identical_rerun": same, "us": round(us, 2),
                          "issued_tflops": round(3 * 2.0 * M * N * K / us / 1e6, 1),
                          "frac_of_tf32_burst": round(3 * 2.0 * M * N * K / us / 1e6 / 819.45, 3)}), flush=True)
    os.environ.pop("JRR_GEMM_PROBE_REPS", None)


if __name__ == "__main__":
    main()
