#!/usr/bin/env python
"""Diagnostic: raw TMA tile-load rate per SM and per chip for the GEMM kernels' box shape (jrr_debug_tma_probe)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402
from jrr_b200 import _lib  # noqa: E402


def main():
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rows, cols = 8192, 1024
    src = torch.randn(rows, cols, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    iters = 2000

    def run(box_rows, boxes, stages, shared, grid, dwell=0):
        def go(n):
            _lib.check(L.jrr_debug_tma_probe(C.c_void_p(src.data_ptr()), rows, cols, box_rows, boxes, stages, shared, n, grid, dwell, st), "probe")
        go(iters); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record(); go(iters); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0.record(); go(1); e1.record(); torch.cuda.synchronize()
        ms = best - e0.elapsed_time(e1)
        byt = iters * boxes * box_rows * 128
        print(json.dumps({"box_rows": box_rows, "boxes": boxes, "stages": stages, "shared": shared, "grid": grid, "dwell_ns": dwell,
                          "ring_KB": boxes * stages * box_rows * 128 // 1024, "B_per_clk_per_SM": round(byt / (ms * 1e-3 * 1.965e9), 1),
                          "chip_TBps": round(byt * grid / (ms * 1e-3) / 1e12, 2), "clk_per_stage": round(ms * 1e-3 * 1.965e9 / iters)}), flush=True)

    for grid in (1, 16, 64, 148):
        for shared in (0, 1):
            for box_rows, boxes, stages in ((128, 3, 4), (128, 4, 3), (128, 1, 12), (128, 2, 6), (128, 6, 2), (128, 12, 1), (64, 6, 4), (256, 2, 3), (128, 3, 2), (128, 3, 1)):
                run(box_rows, boxes, stages, shared, grid)
    run(128, 3, 4, 0, 148, 300)
    run(128, 3, 4, 0, 148, 600)


if __name__ == "__main__":
    main()
