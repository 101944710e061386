#!/usr/bin/env python
"""Diagnostic: where does the A-through-TMEM 3xTF32 GEMM (critic layers) lose its time?  Times jrr_debug_gemm(impl=2)
with the kernel's probe knobs (JRR_GEMM_PROBE bits: 1 = MMAs do not wait for the smem->TMEM A staging, 2 = two MMAs per
k-step, 4 = one MMA per k-step) on shapes that fill whole waves of 148 SMs or not.  Kernel time = (t[11 launches] -
t[1 launch]) / 10, so the operand split kernel of the debug entry point cancels.  Prints JSON lines."""
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
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    shapes = [(4736, 1024, 1024, 148), (4736, 1024, 1024, 74), (4736, 1024, 1024, 37), (4096, 1024, 1024, 128), (4096, 1024, 1024, 148)]
    for M, N, K, grid in shapes:
        A = torch.randn(M, K, device=dev)
        B = torch.randn(N, K, device=dev)
        os.environ["JRR_GEMM_PROBE_GRID"] = str(grid)
        for probe in (0, 1, 9, 15):
            os.environ["JRR_GEMM_PROBE"] = str(probe)
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
            nmma = {0: 3, 1: 3, 3: 2, 7: 1, 9: 3, 15: 1}[probe]
            tiles = (M // 128) * ((N + 127) // 128)
            issued = nmma * 2.0 * M * N * K
            print(json.dumps({"M": M, "N": N, "K": K, "grid": grid, "probe": probe, "mma_per_kstep": nmma, "tiles": tiles,
                              "waves": round(tiles / 148, 2), "us": round(us, 2), "issued_tflops": round(issued / us / 1e6, 1),
                              "frac_of_tf32_burst": round(issued / us / 1e6 / (peaks.get("bf16_tflops", 1638.9) / 2), 3),
                              "clk_per_stage_at_1965": round(us * 1965 / (max(1, -(-tiles // grid)) * (K // 32)), 1)}), flush=True)
    os.environ.pop("JRR_GEMM_PROBE", None)
    os.environ.pop("JRR_GEMM_PROBE_GRID", None)
    # two row blocks per CTA (gemm_ts2_kernel) against the 128x128 kernel, same shapes
    for M, N, K in [(4096, 1024, 1024), (4096, 1024, 768), (4096, 768, 1024), (4736, 1024, 1024), (2048, 1024, 1024), (8192, 1024, 1024)]:
        A = torch.randn(M, K, device=dev)
        B = torch.randn(N, K, device=dev)
        row = {"M": M, "N": N, "K": K}
        for tag, ts1 in (("ts2_us", False), ("ts1_us", True)):
            if ts1:
                os.environ["JRR_GEMM_PROBE_TS1"] = "1"
            else:
                os.environ.pop("JRR_GEMM_PROBE_TS1", None)
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
            row[tag] = round((ts[11] - ts[1]) / 10 * 1e3, 2)
        row["ts2_issued_tflops"] = round(3 * 2.0 * M * N * K / row["ts2_us"] / 1e6, 1)
        row["ts2_frac_of_tf32_burst"] = round(row["ts2_issued_tflops"] / (peaks.get("bf16_tflops", 1638.9) / 2), 3)
        print(json.dumps(row), flush=True)
    os.environ.pop("JRR_GEMM_PROBE_TS1", None)
    os.environ.pop("JRR_GEMM_PROBE_REPS", None)


if __name__ == "__main__":
    main()
