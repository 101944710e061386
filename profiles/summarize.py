#!/usr/bin/env python
"""Turns the raw artefacts of `benchmarks/gpu_validate.sh` (gpurun_out/r1_*) into the committed
summaries under profiles/:  r1_launches_summary.txt (share of device time per kernel from the ncu
launch list) and r1_ncu_top_kernels.txt (selected `ncu --set full` metrics of the top kernels).
Needs `ncu` on PATH (reads the .ncu-rep; no GPU required)."""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("jrr::", "")


def launches():
    rows = [r for r in csv.reader(open(os.path.join(SRC, f"{TAG}_launches.csv"))) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        t = float(r[vi].replace(",", ""))
        t = t / 1000.0 if r[ui] == "ns" else t * 1000.0 if r[ui] in ("ms", "msecond") else t
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(DST, f"{TAG}_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
                + (" --no-secondary" if TAG != "r1" else "") + "\n")
        f.write("(cold-cache, serialised per-launch times: compare SHARES, not absolutes; the command also runs set-up,\n"
                " warm-up, the end-to-end leg, the profiled steps, the other loss-path formulation and the refit, so their kernels\n"
                " appear too: fold_kernel runs once per regressor version -- here at every set_regressor of the set-up and quality\n"
                " checks, never inside a timed step; fused_fwd/fused_bwd/loss_seed belong to the per-vertex formulation)\n\n")
        f.write(f"{'kernel':62s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:62]:62s} {n:8d} {t:10.1f} {t / n:9.2f} {100 * t / tot:6.2f}%\n")
        # shares WITHIN one refinement step of the headline (folded) formulation: one launch of each kernel per step,
        # average per-launch time from the list above -- the number to hold against bench.py's roofline.share_of_step
        step = ["pose_fwd_kernel<2>", "gemm_tc_kernel<128, 0, 0>", "folded_seed_kernel", "gemm_tc_kernel<224, 3, 0>",
                "pose_bwd_kernel<2, 0>", "critic_pre_kernel", "gemm_tc_kernel<128, 1, 1>", "gemm_tc_kernel<128, 4, 1>",
                "critic_head_light_kernel", "gemm_tc_kernel<128, 2, 1>", "gemm_tc_kernel<96, 3, 1>", "critic_post_kernel",
                "loss_finish_kernel", "adam_coef_kernel", "adam_params_kernel", "bump_step_kernel"]
        grp = {"critic_gemm_fwd": ["gemm_tc_kernel<128, 1, 1>", "gemm_tc_kernel<128, 4, 1>"],
               "critic_gemm_bwd": ["gemm_tc_kernel<128, 2, 1>", "gemm_tc_kernel<96, 3, 1>"]}
        if any(k.startswith("gemm_pair_kernel") for k in agg):      # round 2: every GEMM of the step on CTA pairs
            step = ["pose_fwd_kernel<2>", "gemm_pair_kernel<256, 0>", "folded_seed_kernel", "gemm_pair_kernel<224, 3>",
                    "pose_bwd_kernel<2, 0>", "critic_pre_kernel", "gemm_pair_kernel<256, 1>", "gemm_pair_kernel<256, 4>",
                    "gemm_pair_kernel<256, 2>", "gemm_pair_kernel<192, 3>", "critic_post_kernel", "loss_finish_kernel",
                    "adam_coef_kernel", "adam_params_kernel"]
            grp = {"critic_gemm_fwd": ["gemm_pair_kernel<256, 1>", "gemm_pair_kernel<256, 4>"],
                   "critic_gemm_bwd": ["gemm_pair_kernel<256, 2>", "gemm_pair_kernel<192, 3>"]}
        have = [(k, agg[k][1] / agg[k][0]) for k in step if k in agg]
        if have:
            st = sum(t for _, t in have)
            f.write(f"\none folded refinement step = {len(have)} launches, {st:.1f} us serialised under ncu\n")
            for k, t in sorted(have, key=lambda kv: -kv[1]):
                f.write(f"{k[:62]:62s} {1:8d} {t:10.1f} {t:9.2f} {100 * t / st:6.2f}%\n")
            d = dict(have)
            for g, ks in grp.items():
                if all(k in d for k in ks):
                    f.write(f"{g + ' (bench.py kernel group)':62s} {len(ks):8d} {sum(d[k] for k in ks):10.1f} {'':9s} {100 * sum(d[k] for k in ks) / st:6.2f}%\n")


def top_kernels():
    import json
    out = ["ncu --set full --clock-control none --import-source on -k regex:'folded_seed|gemm_tc_kernel' -s 42 -c 14 "
           "python bench.py --steps 3 --warmup 3 --no-cpu-baseline" if TAG == "r1" else
           "ncu --set full --clock-control none --import-source on -k regex:'gemm_pair|folded_seed|critic_pre|critic_post|pose_fwd|"
           "pose_bwd|adam_params' -s 60 -c 26 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary",
           "ncu --set full ... -k regex:'fused_bwd|fused_fwd' -s 30 -c 4 python bench.py --loss-path vertex --steps 3 --warmup 3 --no-cpu-baseline",
           "ncu --set full ... -k regex:'fused_fwd|fused_bwd|smpl_small' -c 14 python benchmarks/module_calls.py   (module path: SMPL.forward / backward at 4096, 256, 8 poses)",
           "ncu --set full ... -k regex:'sil_|small_bwd' -c 12 python benchmarks/sil_calls.py   (one refinement iteration with the silhouette term, 1024 frames at 224 x 224)",
           "ncu --set full ... -k regex:'small_bwd|smpl_small' -s 0 -c 6 python benchmarks/module_calls.py with B in (8, 1) only   (single-launch small-batch paths)",
           "(first captured launch of each kernel; B = 4096 poses, dense 17x6890 regressor; times under ncu are cold-cache and serialised)", ""]
    traffic = {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for tag in (f"{TAG}_prof", f"{TAG}_prof_vertex", f"{TAG}_prof_module", f"{TAG}_prof_sil", f"{TAG}_prof_smallbwd"):
        rep = os.path.join(SRC, tag + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        seen = set()
        for r in rows[2:]:
            name = short(r[ki])
            if name in seen:
                continue
            seen.add(name)
            out.append(f"== {name}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    out.append(f"   {w:95s} {r[i]:>18s} {units[i]}")
            out.append("")
            # per-launch DRAM traffic of each captured kernel, read by bench.py for roofline.traffic
            traffic[name] = float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
    open(os.path.join(DST, f"{TAG}_ncu_top_kernels.txt"), "w").write("\n".join(out))
    json.dump({"source": f"profiles/{TAG}_ncu_top_kernels.txt (ncu --set full, B = 4096, dense regressor)",
               "dram_bytes_per_launch": traffic}, open(os.path.join(DST, "ncu_traffic.json"), "w"), indent=1)


def main():
    for f in (f"{TAG}_bench.json", f"{TAG}_bench_k20.json", f"{TAG}_bench_vertex.json", f"{TAG}_bench_shipped.json", f"{TAG}_bench_reference.json", f"{TAG}_gpu_tests.txt",
              f"{TAG}_sweep.jsonl", f"{TAG}_memcheck.txt", f"{TAG}_bench_2gpu.json", f"{TAG}_bench_8gpu.json", f"{TAG}_launches.csv",
              f"{TAG}_gemm_pair_check.jsonl", f"{TAG}_gemm_role_stamps.jsonl", f"{TAG}_step_breakdown.jsonl",
              f"{TAG}_multi_gpu_check_2.json", f"{TAG}_multi_gpu_check_8.json", f"{TAG}_module_launches.csv", f"{TAG}_silhouette_launches.csv",
              f"{TAG}_smoke.txt"):
        p = os.path.join(SRC, f)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST, f))
    launches()
    top_kernels()
    print("profiles/ refreshed for", TAG)


if __name__ == "__main__":
    main()
