#!/usr/bin/env python
"""Headline benchmark: refine-step poses/sec (SMPL fwd + 17x6890 J-regressor + loss + analytic
bwd + Adam) on config C2 of BASELINE.json -- 4096 synthetic frames per GPU, one "step" = one
Adam iteration of the refinement loop over the whole batch (optimize.py:220-265).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); frames shard with no data-path
collective (weak scaling: every rank refines its own 4096 frames).  Rank 0 prints ONE JSON
line.  `--impl reference` times the CPU oracle port of the reference path on the host cores.

What the line carries (every leg goes through the package's public API):
  value      K iterations of `PoseRefiner.refine` on device-resident frames (CUDA graphs of 10 iterations)
  e2e        the same K iterations from PINNED HOST buffers: H2D of x6/betas/gt, every iteration's loss
             copied back, one regressor update (`RegressorRefit.step`: accumulate, all-reduce when N > 1,
             Adam, re-fold of the loss-path operator) per <= 100 iterations, D2H of the refined parameters
  roofline   dominant kernel group, CUDA-event timed in isolation -> against the BURST tensor peak;
             whole_step -> against the SUSTAINED peak
  quality    MPJPE before/after + `oracle_one_step`: this run's first iteration (losses, gradient) against the
             fp64 CPU oracle on the same 4096 frames
  secondary  C3 (312 000 frames x 100 iterations, STRONG scaling, the whole per-batch loop with the critic and
             regressor all-reduces), C4 (refit time split), C5 (SMPL module forward / backward sweep)
  cpu_baseline / gpu_eager_reference   the oracle port on the host cores / on the same B200 in eager torch
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "refine_step_poses_per_sec"
UNIT = "poses/s"
FRAMES = 4096
C3_FRAMES = 312000
WORKLOAD = ("C2: optimize.py pose refinement, 4096 synthetic frames/GPU x Adam iterations, "
            "loss 10000*joint_MSE + 10*pose_critic_MSE, random-init SMPL (6890 v, 24 j, 10 betas, 207 pose dims)")

# algorithmic work per pose-step ("useful-minimum" formulation)
F_USEFUL = 28.5e6
F_GEMM = 24.46e6                        # tensor-eligible (blend + critic, fwd + bwd)
F_CRITIC_DIR = 2.0 * (768 * 1024 + 1024 * 1024)      # one direction of the critic's two wide layers
ALG_BYTES_PER_POSE_STEP = 3900
SMPL_FWD_BYTES, SMPL_BWD_BYTES = 84172, 85076        # SURVEY.md 8d, module path
SMPL_FWD_GEMM_FLOPS = 2.0 * 20670 * 218              # blend contraction per pose, one direction


_REAL_STDOUT = None


def quiet_stdout():
    """Library chatter (e.g. NCCL's version banner) goes to stderr: stdout carries exactly one
    JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------- inputs
def load_regressor(kind):
    """dense: |N(0,1)| everywhere (the full 17x6890 reduction); shipped: the reference artefact through the
    product loader -- from the reference tree when present, else its byte-for-byte copy under tests/golden/."""
    import torch
    import jrr_b200 as jrr
    if kind == "dense":
        return torch.from_numpy(jrr.synthetic.make_dense_regressor(0))
    for p in (os.path.join(os.environ.get("JRR_REFERENCE_ROOT", "/root/reference"), "models", "retrained_J_Regressor.pt"),
              os.path.join(ROOT, "tests", "golden", "retrained_J_Regressor.pt")):
        if os.path.exists(p):
            return jrr.load_j_regressor(p)
    raise SystemExit("retrained_J_Regressor.pt not found")


def make_problem(jrr, smpl, J, n, seed, dev, chunk=16384):
    """Synthetic frames on the device; GT joints come from the CUDA path itself."""
    import torch
    inp = jrr.synthetic.make_pose_inputs(n, seed)
    R = torch.from_numpy(inp["true_rotmat"]).to(dev)
    tb = torch.from_numpy(inp["true_betas"]).to(dev)
    gt = torch.empty(n, 17, 3, device=dev)
    Jd = J.to(dev)
    with torch.no_grad():
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            gt[lo:hi] = 1000 * jrr.move_pelvis(jrr.find_joints(smpl, tb[lo:hi], R[lo:hi, :1], R[lo:hi, 1:], Jd))
    gt += torch.from_numpy(inp["gt_noise"]).to(dev)
    return torch.from_numpy(inp["x6"]).to(dev).contiguous(), torch.from_numpy(inp["betas"]).to(dev).contiguous(), gt.contiguous()


def mpjpe_of(jrr, smpl, J, x6, be, gt, dev, chunk=16384):
    import torch
    tot = pa = 0.0
    n = x6.shape[0]
    Jd = J.to(dev)
    with torch.no_grad():
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            Rg = jrr.rot6d_to_rotmat(x6[lo:hi].reshape(-1, 6)).view(-1, 24, 3, 3)
            m, p = jrr.evaluate(jrr.find_joints(smpl, be[lo:hi], Rg[:, :1], Rg[:, 1:], Jd), gt[lo:hi])
            tot += float(m) * (hi - lo)
            pa += float(p) * (hi - lo)
    return tot / n, pa / n


# ---------------------------------------------------------------------------------------- oracle legs (baselines / checker)
def oracle_problem(O, jrr, J, n, seed, device="cpu", dtype=None, chunk=1024):
    import torch
    dtype = dtype or torch.float32
    model = jrr.synthetic.make_smpl_model(0)
    osmpl = O.OracleSMPL(model, dtype, device)
    inp = jrr.synthetic.make_pose_inputs(n, seed)
    t = {k: torch.from_numpy(v).to(device=device, dtype=dtype) for k, v in inp.items()}
    Jd = J.to(device=device, dtype=dtype)
    gt = torch.cat([O.make_gt(osmpl, Jd, t["true_rotmat"][lo:lo + chunk], t["true_betas"][lo:lo + chunk], t["gt_noise"][lo:lo + chunk])
                    for lo in range(0, n, chunk)])
    return osmpl, Jd, t["x6"], t["betas"], gt


def oracle_rate(J, n_frames, iters, warm, device="cpu", seed=0, chunk=1024, regressor_updates=True):
    """The oracle port of the reference path (eager torch; `device` = cpu with all host threads, or the B200):
    pose-steps/s of the loop of optimize.py:220-265 (+ one regressor update per <= 100 iterations, :300-312).
    The batch is processed in chunks of `chunk` frames with logical_batch = n_frames (identical arithmetic, bounded
    memory: the reference's per-vertex transforms alone are 1.8 GB at 4096 frames)."""
    import torch
    import jrr_b200 as jrr
    from oracle import jrr_oracle as O
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    osmpl, Jd, x6, be, gt = oracle_problem(O, jrr, J, n_frames, seed, device, chunk=chunk)
    sd = {k: v.to(device) for k, v in O.make_critic_state_dict(0).items()}
    x6 = x6.clone().requires_grad_(True)
    be = be.clone().requires_grad_(True)
    opt = torch.optim.Adam([x6, be], lr=1e-2)
    radam = O.RegressorAdam(Jd, lr=1e-2)

    def step():
        opt.zero_grad()
        for lo in range(0, n_frames, chunk):
            sl = slice(lo, min(n_frames, lo + chunk))
            total, _, _, _ = O.refine_loss(osmpl, radam.J.detach(), sd, x6[sl], be[sl], gt[sl], logical_batch=n_frames)
            total.backward()
        opt.step()

    def regressor_update():
        g = None
        for lo in range(0, n_frames, chunk):
            sl = slice(lo, min(n_frames, lo + chunk))
            gi, _ = O.regressor_grad(osmpl, radam.J.detach(), x6.detach()[sl], be.detach()[sl], gt[sl], logical_batch=n_frames)
            g = gi if g is None else g + gi
        radam.step(g)

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()
    for _ in range(warm):
        step()
    sync()
    t0 = time.perf_counter()
    for i in range(iters):
        step()
        if regressor_updates and ((i + 1) % 100 == 0 or i == iters - 1):
            regressor_update()
    sync()
    dt = time.perf_counter() - t0
    return n_frames * iters / dt, dt / iters * 1e3, cores


def oracle_one_step(J, x6, betas, gt, chunk=512):
    """CHECKER: fp64 CPU oracle losses and gradient [n,154] of one refinement iteration on these frames."""
    import torch
    import jrr_b200 as jrr
    from oracle import jrr_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    n = x6.shape[0]
    osmpl = O.OracleSMPL(jrr.synthetic.make_smpl_model(0), torch.float64)
    sd = {k: v.double() for k, v in O.make_critic_state_dict(0).items()}
    tot = jl = pl = 0.0
    grads = []
    for lo in range(0, n, chunk):
        x = x6[lo:lo + chunk].double().requires_grad_(True)
        b = betas[lo:lo + chunk].double().requires_grad_(True)
        t, j, p, _ = O.refine_loss(osmpl, J.double(), sd, x, b, gt[lo:lo + chunk].double(), logical_batch=n)
        t.backward()
        tot, jl, pl = tot + t.item(), jl + j.item(), pl + p.item()
        grads.append(torch.cat([x.grad.reshape(-1, 144), b.grad], dim=1))
    kink = O.critic_kink_frames(sd, x6)
    return (tot, jl, pl), torch.cat(grads), kink


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.frames
    J = load_regressor(args.regressor)
    rate, ms, cores = oracle_rate(J, n, args.steps, args.warmup)
    sample = (f"{n} frames per step (the CUDA arm's batch), {args.regressor} regressor, {args.steps} Adam iterations + one "
              f"regressor update per <= 100 iterations, oracle port in eager torch, {cores} threads, fp32")
    emit({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": n, "regressor": args.regressor},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ---------------------------------------------------------------------------------------- helpers of the CUDA arm
class Timer:
    """CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world):
        import torch
        self.torch, self.dev, self.world = torch, dev, world
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def _fence(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn):
        torch = self.torch
        self._fence()
        self.e0.record()
        fn()
        self.e1.record()
        self._fence()
        t = torch.tensor([self.e0.elapsed_time(self.e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return t.item()


def median_ms(fn, torch, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def kernel_table(acc, folded, B, pk, l2_bwd_products=3):
    """Per kernel-group times of jrr_refine_step_profiled -> roofline entries.  Groups are timed in isolation
    (events between serialised launches), so tensor-bound ones are quoted against the BURST peak."""
    tf32_burst = pk["bf16_burst"] / 2
    kern = []
    fused_fwd = not folded and acc.get("skin_fwd", 0.0) < 0.01
    fused_bwd = not folded and acc.get("blend_gemm_bwd", 0.0) < 0.01
    rename = {"blend_gemm_fwd": "folded_gemm_fwd(Q=feat.T^T,N=1224)", "blend_gemm_bwd": "folded_gemm_bwd(dfeat=dQ.T,K=1224)",
              "loss_seed": "folded_seed(joints+loss+dA+dQ)"}
    for name, ms in acc.items():
        if ms <= 0 or (fused_fwd and name == "skin_fwd") or (fused_bwd and name == "blend_gemm_bwd"):
            continue
        if name == "critic_head" and ms < 0.004:
            continue                                     # head-less chain: an empty event interval
        if folded and name in ("skin_fwd", "skin_bwd", "dA_reduce"):
            continue                                     # empty event intervals on the folded path
        e = {"name": name, "ms": round(ms, 4)}
        fl = by = None
        if folded and name in rename:
            e["name"] = rename[name]
            if name != "loss_seed":
                fl = 3 * 2.0 * B * 1224 * 218
        elif name == "blend_gemm_fwd":
            fl = 3 * 2.0 * B * 20670 * 218
            if fused_fwd:
                e["name"] = "fused_fwd(blend_gemm+skinning+regressor)"
        elif name == "blend_gemm_bwd" or (name == "skin_bwd" and fused_bwd):
            fl = 3 * 2.0 * B * 20670 * 217
            if fused_bwd:
                e["name"] = "fused_bwd(skinning_bwd+blend_gemm_bwd)"
        elif name in ("critic_gemm_fwd", "critic_gemm_bwd", "critic_chain"):
            fl = 3 * B * F_CRITIC_DIR * (2 if name == "critic_chain" else 1)
            if name != "critic_gemm_fwd":        # the layer-2 backward may issue two products per K step (0/1 operand)
                fl -= (3 - l2_bwd_products) * B * 2.0 * 1024 * 1024
        elif name == "skin_fwd":
            by = 4.0 * B * (20670 + 288 + 51)
        elif name == "skin_bwd":
            by = 4.0 * B * (20670 + 288 + 51 + 2 * 20670 + 288)
        if fl is not None:
            e.update(bound="tensor", achieved=round(fl / (ms * 1e-3) / 1e12, 2), peak=tf32_burst, unit="TFLOP/s")
        elif by is not None:
            e.update(bound="hbm", achieved=round(by / (ms * 1e-3) / 1e9, 2), peak=pk["hbm_gbs"], unit="GB/s")
        if "achieved" in e:
            e["frac"] = round(e["achieved"] / e["peak"], 4)
        kern.append(e)
    kern.sort(key=lambda e: -e["ms"])
    return kern


def balanced_chunks(n, max_chunk):
    """[lo, hi) bounds of ceil(n / max_chunk) chunks whose sizes differ by at most one frame."""
    k = max(1, math.ceil(n / max_chunk))
    return [((n * i) // k, (n * (i + 1)) // k) for i in range(k)]


def run_c3(jrr, smpl, J, sd, args, dev, rank, world, timer):
    """C3 / C4: 312 000 frames x 100 iterations, STRONG scaling.  Every rank owns a contiguous range, cut into the same
    number of balanced chunks (<= 4096 frames, sizes differing by at most one, so at most two graph captures); chunk b of
    all ranks together is global batch b: refine 100 iterations -> critic training step (7.36 MB all-reduce) -> regressor
    refit step (468 524 B all-reduce), i.e. `RefinementLoop.run_batch` -- the loop of optimize.py:144-312 with its real
    exchange steps inside the timed region."""
    import torch
    lo, hi = jrr.shard_range(args.c3_frames, rank, world)
    n = hi - lo
    x6, be, gt = make_problem(jrr, smpl, J, n, 1000 + rank, dev)
    n_batches = max(math.ceil((jrr.shard_range(args.c3_frames, r, world)[1] - jrr.shard_range(args.c3_frames, r, world)[0]) / args.frames)
                    for r in range(world))
    def bounds(nr):
        return [((nr * i) // n_batches, (nr * (i + 1)) // n_batches) for i in range(n_batches)]
    mine = bounds(n)
    sizes = [bounds(jrr.shard_range(args.c3_frames, r, world)[1] - jrr.shard_range(args.c3_frames, r, world)[0]) for r in range(world)]
    LB = [sum(s[b][1] - s[b][0] for s in sizes) for b in range(n_batches)]
    loop = jrr.RefinementLoop(smpl, J, sd, None, refine_iters=args.c3_iters, chunk=args.frames, loss_path=args.loss_path)
    mp0, _ = mpjpe_of(jrr, smpl, loop.refit.J_regressor, x6, be, gt, dev)
    # warm-up: capture the graphs of both chunk sizes on scratch copies (the critic / regressor state this touches is
    # part of the synthetic set-up, not of the measurement)
    for sz in (max(b - a for a, b in mine), min(b - a for a, b in mine)):      # always two: same collective count on every rank
        loop.run_batch({"orient": x6[:sz, :1].clone(), "pose": x6[:sz, 1:].clone(), "betas": be[:sz].clone(), "gt_j3d": gt[:sz]},
                       global_batch=sz * world)
    out_x6, out_be = torch.empty_like(x6), torch.empty_like(be)

    def body():
        for b, (a, c) in enumerate(mine):
            o = loop.run_batch({"orient": x6[a:c, :1], "pose": x6[a:c, 1:], "betas": be[a:c], "gt_j3d": gt[a:c]}, global_batch=LB[b])
            out_x6[a:c] = o["x6"]
            out_be[a:c] = o["betas"]
    ms = timer.run(body)
    mp1, _ = mpjpe_of(jrr, smpl, loop.refit.J_regressor, out_x6, out_be, gt, dev)
    c3 = {"frames": args.c3_frames, "iterations": args.c3_iters, "n_gpus": world, "scaling": "strong",
          "global_batches": n_batches, "frames_per_rank_per_batch": sorted({b - a for a, b in mine}),
          "seconds": round(ms / 1e3, 4), "pose_steps_per_s": round(args.c3_frames * args.c3_iters / (ms / 1e3), 1),
          "collectives_per_batch": "all-reduce of the critic gradient (7 360 612 B) + loss, all-reduce of the regressor "
                                   "gradient (468 520 B) + loss" if world > 1 else "none (one rank)",
          "mpjpe_before_mm_rank0": round(mp0, 3), "mpjpe_after_mm_rank0": round(mp1, 3)}
    # C4: one refit step over ALL frames of the set (accumulate over this rank's range / all-reduce / apply)
    refit = loop.refit
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    res = None
    for rep in range(2):
        refit.G.zero_(); refit.loss.zero_()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ev[0].record()
        refit.accumulate(out_x6, out_be, gt, logical_batch=args.c3_frames)
        ev[1].record()
        if world > 1:
            torch.distributed.all_reduce(refit.G); torch.distributed.all_reduce(refit.loss)
        ev[2].record()
        refit.native.regressor_apply(refit.J, refit.mask, refit.G, refit.m, refit.v, refit.t, refit.lr)
        ev[3].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(3)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        res = t.tolist()
    c4 = {"frames": args.c3_frames, "n_gpus": world, "accumulate_ms": round(res[0], 3), "allreduce_ms": round(res[1], 3),
          "apply_ms": round(res[2], 3), "allreduce_bytes": 17 * 6890 * 4 + 4,
          "accumulate_gbs_vs_fused_820B_per_frame": round(820.0 * n / (res[0] * 1e-3) / 1e9, 2)}
    del loop
    return c3, c4


def run_silhouette(jrr, smpl, J, sd, dev, n=1024, S=224):
    """Widening row 8f-4: the silhouette term at the reference's image size (optimize.py:111: 224) -- rasteriser forward and
    backward alone, and one whole refinement iteration with every term of optimize.py:252-253 (3-D joints, pose critic,
    2-D reprojection, silhouette) through PoseRefiner.refine_silhouette."""
    import torch
    inp = jrr.synthetic.make_pose_inputs(n, 5)
    R = torch.from_numpy(inp["true_rotmat"]).to(dev)
    # (zero shape coefficients: the synthetic shape directions are spatially UNCORRELATED noise, which with |beta| ~ 1 tears
    # neighbouring vertices 6 cm apart -- faces of 25-50 pixels in their bounding box instead of SMPL's 4-9)
    betas = torch.zeros(n, 10, device=dev)
    faces = jrr.synthetic.make_local_faces(smpl._model_np["v_template"], lbs_weights=smpl._model_np["lbs_weights"])
    rend = jrr.Mesh_Renderer(image_size=S, faces=faces)
    cam = torch.tensor([0.0, 0.4, 5000.0 / S * 2.3], device=dev).repeat(n, 1).contiguous()
    verts = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False).vertices.contiguous()
    from jrr_b200.mesh_renderer import _backward, _forward
    mesh = rend.mesh(6890, dev)
    # targets of a real batch: the silhouette and the joints of the TRUE poses (the refinement converges towards them, so the
    # mesh -- and with it the rasteriser's work -- stays what it is over the timed iterations)
    alpha, p2f, _ = _forward(mesh, verts, cam, S, True)
    mask = (alpha > 0.5).float().unsqueeze(1)
    tgt = mask[:, 0].contiguous()
    f_ms = median_ms(lambda: _forward(mesh, verts, cam, S, True, tgt, n), torch, reps=10)
    b_ms = median_ms(lambda: _backward(mesh, verts, cam, S, True, alpha, p2f, target=tgt, logical_batch=n, weight=100.0), torch, reps=10)
    ref = jrr.PoseRefiner(smpl, J, sd, chunk=n, use_graph=False)
    x6 = torch.from_numpy(inp["x6"]).to(dev).reshape(n, 24, 6).contiguous()
    with torch.no_grad():
        j17 = jrr.find_joints(smpl, betas, R[:, :1], R[:, 1:], J.to(dev) if hasattr(J, "to") else J)
    gt = (1000.0 * jrr.move_pelvis(j17)).contiguous()
    view = j17 * torch.tensor([-2.0, -2.0, 2.0], device=dev) + cam[:, None, :]
    gt2d = ((224 - 1.0) / 2.0 * (1.0 - (5000.0 / 224.0) * view[..., :2] / view[..., 2:3])).contiguous()      # renderer.py:35-49

    def run(iters):
        xw, bw, cw = x6.clone(), betas.clone(), cam.clone()
        ref.refine_silhouette(xw, bw, cw, gt, gt2d, mask, rend, iters=iters)

    it12 = median_ms(lambda: run(12), torch, warm=1, reps=3)
    it4 = median_ms(lambda: run(4), torch, warm=1, reps=3)
    it_ms = (it12 - it4) / 8            # one iteration (the per-call copies cancel out)
    covered = (p2f >= 0).float().mean().item()
    del ref
    return {"frames": n, "image_size": S, "faces": int(faces.shape[0]), "covered_pixel_fraction": round(covered, 4),
            "raster_fwd_ms": round(f_ms, 3), "raster_bwd_ms": round(b_ms, 3),
            "refine_iteration_all_terms_ms": round(it_ms, 3),
            "frames_per_s_all_terms": round(n / (it_ms * 1e-3)),
            "note": "synthetic triangle soup (2 local triangles per vertex, zero betas); the iteration = module forward + rasteriser + "
                    "rasteriser backward + module backward + the fused refinement step with the 2-D term"}


def run_c5(jrr, smpl, dev, pk, sizes):
    """C5: the drop-in module itself -- SMPL forward (vertices + 49 joints out) and forward+backward, rotation-matrix
    inputs, through NativeModel.smpl_forward / smpl_backward (what SMPLFunction calls)."""
    import torch
    nat = smpl.native()
    tf32_burst = pk["bf16_burst"] / 2
    rows = []
    base = jrr.synthetic.make_pose_inputs(4096, 7)
    for B in sizes:
        rep = (B + 4095) // 4096
        full = torch.from_numpy(base["true_rotmat"]).repeat(rep, 1, 1, 1)[:B].reshape(B, 24, 9).to(dev).contiguous()
        b = torch.from_numpy(base["true_betas"]).repeat(rep, 1)[:B].to(dev).contiguous()
        reps = 20 if B <= 4096 else 5
        f_ms = median_ms(lambda: nat.smpl_forward(b, full, 0, True, True), torch, reps=reps)
        dv = torch.randn(B, 6890, 3, device=dev)
        dj = torch.randn(B, 49, 3, device=dev)
        b_ms = median_ms(lambda: nat.smpl_backward(b, full, 0, dv, dj), torch, reps=reps)
        rows.append({"B": B, "fwd_us": round(f_ms * 1e3, 1), "fwdbwd_us": round((f_ms + b_ms) * 1e3, 1),
                     "fwd_poses_per_s": round(B / f_ms * 1e3),
                     "fwd_hbm_frac": round(SMPL_FWD_BYTES * B / (f_ms * 1e-3) / (pk["hbm_gbs"] * 1e9), 4),
                     "fwd_tensor_frac_3xtf32": round(3 * SMPL_FWD_GEMM_FLOPS * B / (f_ms * 1e-3) / (tf32_burst * 1e12), 4),
                     "bwd_hbm_frac": round(SMPL_BWD_BYTES * B / (b_ms * 1e-3) / (pk["hbm_gbs"] * 1e9), 4)})
        del full, b, dv, dj
    nat._ws, nat._ws_B = None, 0          # the 65 536-pose workspace is tens of GB: hand it back
    nat.ws_generation += 1
    torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--regressor", default="dense", choices=["dense", "shipped"],
                    help="dense 17x6890 (headline: the full reduction) or the shipped sparse artefact")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 / C4 / C5 legs")
    ap.add_argument("--c3-frames", type=int, default=C3_FRAMES)
    ap.add_argument("--c3-iters", type=int, default=100)
    ap.add_argument("--gemm-impl", type=int, default=0)
    ap.add_argument("--steps-per-graph", type=int, default=10,
                    help="Adam iterations captured per CUDA graph in the device-resident leg (the e2e leg reads the loss "
                         "back after every step and replays a one-step graph)")
    ap.add_argument("--loss-path", default="folded", choices=["vertex", "folded"],
                    help="vertex: per-vertex fused kernels (blend GEMM + skinning + 17x6890 reduction, the path SURVEY.md 8a "
                         "names); folded: regressor o skinning o blend operator folded per regressor version (include/jrr.h)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import jrr_b200 as jrr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    K, B = args.steps, args.frames
    U = max(1, min(args.steps_per_graph, K))
    W = max(3, args.warmup, U + 1)            # the warm-up replays both captured graphs (U-step and 1-step)
    timer = Timer(dev, world)
    pk = peaks()

    # identical model / regressor / critic on every rank; frames seeded per rank
    model = jrr.synthetic.make_smpl_model(0)
    smpl = jrr.SMPL(model_dict=model, create_transl=False, gemm_impl=args.gemm_impl).to(dev)
    J0 = load_regressor(args.regressor)
    torch.manual_seed(0)
    sd = jrr.Discriminator().state_dict()             # default init, frozen
    # one regressor tensor on the device, owned by the refit and shared by every refiner below (as in RefinementLoop):
    # the native model holds ONE normalised copy, and objects that work on the same tensor never displace each other
    refit = jrr.RegressorRefit(smpl, J0, lr=1e-2, chunk=B)
    J = refit.J_regressor
    refiner = jrr.PoseRefiner(smpl, J, sd, lr=1e-2, w_joint=10000.0, w_pose=10.0, chunk=B, use_graph=True,
                              loss_path=args.loss_path, steps_per_graph=U)
    folded = args.loss_path == "folded"
    x6_0, be_0, gt = make_problem(jrr, smpl, J, B, seed=rank, dev=dev)
    x6_pin, be_pin, gt_pin = x6_0.cpu().pin_memory(), be_0.cpu().pin_memory(), gt.cpu().pin_memory()

    # ---------------------------------------------------------------- first iteration against the fp64 oracle (checker)
    one_step = None
    eager = jrr.PoseRefiner(smpl, J, sd, chunk=B, use_graph=False, loss_path=args.loss_path)
    st = eager._buffers(B)
    st["x6"].copy_(x6_0); st["betas"].copy_(be_0); st["gt"].copy_(gt)
    eager._run_chunk(st, 1, B)
    if rank == 0 and not args.no_cpu_baseline:
        (ot, oj, op_), og, kink = oracle_one_step(J0, x6_0.cpu(), be_0.cpu(), gt.cpu())
        m10 = st["m"].cpu().double() * 10            # Adam's first moment after one step from zero state = 0.1 * gradient
        ls = st["loss"].cpu().double()
        dfr = (m10 - og).abs().max(1).values / og.abs().max()
        one_step = {"loss_rel_err": abs(ls[0].item() - ot) / ot, "joint_loss_rel_err": abs(ls[1].item() - oj) / oj,
                    "pose_loss_rel_err": abs(ls[2].item() - op_) / op_,
                    "gradient_rel_err": dfr[~kink].max().item(),
                    "relu_kink_frames": int(kink.sum()),
                    "gradient_rel_err_on_kink_frames": dfr[kink].max().item() if kink.any() else 0.0,
                    "frames": B, "against": "fp64 CPU oracle (oracle/jrr_oracle.py), same inputs, tolerance 1e-5 (losses) / 1e-4 "
                                            "(gradient, max-abs difference over max-abs gradient); frames with a critic pre-activation "
                                            "within fp32 round-off of a ReLU kink (oracle.critic_kink_frames) are reported separately"}
    del eager

    # ---------------------------------------------------------------- device-resident timing (`value`)
    xw, bw = x6_0.clone(), be_0.clone()
    refiner.refine(xw, bw, gt, iters=W)               # warm-up: captures the 1-step and the U-step graph
    x6, be = x6_0.clone(), be_0.clone()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timer.run(lambda: refiner.refine(x6, be, gt, iters=K))
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms_total * 1e-3)
    main_launches = refiner.launches_per_step
    mpjpe, pampjpe = mpjpe_of(jrr, smpl, J, x6, be, gt, dev)
    mp0, _ = mpjpe_of(jrr, smpl, J, x6_0, be_0, gt, dev)

    # ---------------------------------------------------------------- end to end (host buffers, public API)
    loss_pin = torch.zeros(K, 5).pin_memory()
    work_x, work_b = torch.empty_like(x6_pin).pin_memory(), torch.empty_like(be_pin).pin_memory()
    n_refits = math.ceil(K / 100)

    dx, db, dg = torch.empty_like(x6_0), torch.empty_like(be_0), torch.empty_like(gt)     # the batch on the device

    def e2e_pass():
        # One batch of the reference loop from HOST buffers: H2D of the frames (pinned -> device, once per batch: the 100
        # iterations and the regressor update of optimize.py:220-312 all work on the same device-resident batch),
        # PoseRefiner.refine for <= 100 iterations with every iteration's loss copied back to pinned host memory,
        # RegressorRefit.step on the refined frames (accumulate, all-reduce over the ranks, Adam, operator re-fold),
        # D2H of the refined parameters into the pinned buffers
        for i in range(n_refits):
            it = min(100, K - 100 * i)
            dx.copy_(work_x, non_blocking=True); db.copy_(work_b, non_blocking=True); dg.copy_(gt_pin, non_blocking=True)
            refiner.refine(dx, db, dg, iters=it, logical_batch=B, loss_history=loss_pin[100 * i:100 * i + it])
            refit.step(dx, db, dg, logical_batch=world * B)
            work_x.copy_(dx, non_blocking=True); work_b.copy_(db, non_blocking=True)
    work_x.copy_(x6_pin); work_b.copy_(be_pin)
    e2e_pass()                                        # warm-up (the refit's kernels, pinned-copy paths)
    refit.reset(J0)
    refiner.refine(xw, bw, gt, iters=W)               # re-capture both graphs outside the timed region
    work_x.copy_(x6_pin); work_b.copy_(be_pin)
    e2e_ms = timer.run(e2e_pass)
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    h2d = (x6_pin.numel() + be_pin.numel() + gt_pin.numel()) * 4 * n_refits
    d2h = (work_x.numel() + work_b.numel()) * 4 * n_refits + K * 5 * 4
    e2e_loss_last = loss_pin[K - 1].tolist()
    refit.reset(J0)

    # ---------------------------------------------------------------- per-kernel timing / roofline
    P = 5
    acc = {}
    st = refiner._buffers(B)
    st["x6"].copy_(x6_0); st["betas"].copy_(be_0); st["gt"].copy_(gt)
    st["m"].zero_(); st["v"].zero_(); st["t"].zero_()
    for i in range(P + 2):
        ms = refiner.native.refine_step_profiled(st["x6"], st["betas"], st["gt"], st["m"], st["v"], st["t"],
                                                 refiner.lr, refiner.w_joint, refiner.w_pose, logical_batch=B,
                                                 loss_out=st["loss"])
        if i >= 2:
            for k, v in ms.items():
                acc[k] = acc.get(k, 0.0) + v / P
    l2p = refiner.native.critic_layer2_bwd_products(B)
    kern = kernel_table(acc, folded, B, pk, l2p)
    step_ms_prof = sum(e["ms"] for e in kern)
    dom = next((e for e in kern if "achieved" in e), None)
    roofline = None
    if dom is not None:
        roofline = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": None, "kernel": dom["name"],
                    "share_of_step": round(dom["ms"] / step_ms_prof, 3),
                    "frac_of_sustained_peak": round(dom["achieved"] / (pk["bf16_sustained"] / 2), 4) if dom["bound"] == "tensor" else None,
                    "peak_source": pk["source"] + (" bf16 burst / 2 (dense TF32; the group is event-timed in isolation)"
                                                   if dom["bound"] == "tensor" else " hbm copy"),
                    "traffic_note": "not measured in this run (needs ncu); the last ncu --set full capture is profiles/ncu_traffic.json"}
    tf32_sus = pk["bf16_sustained"] / 2               # the whole step runs inside a long replay: sustained peak
    pose_steps_per_s = value / world
    f_gemm = F_GEMM if not folded else 2.0 * (2 * 1224 * 218) + 2 * F_CRITIC_DIR
    issued = 3 * f_gemm - (3 - l2p) * 2.0 * 1024 * 1024           # tensor-core FLOPs actually issued per pose-step
    whole = {"tensor_frac_3xtf32": round(issued * pose_steps_per_s / (tf32_sus * 1e12), 4),
             "tensor_frac_3xtf32_vs_burst": round(issued * pose_steps_per_s / (pk["bf16_burst"] / 2 * 1e12), 4),
             "issued_gflop_per_step": round(issued * B / 1e9, 2),
             # the same results through the generic three-product scheme everywhere (what round 1 issued): the step's
             # rate of 3xTF32-equivalent work -- not a utilisation figure, the issued one above is
             "tensor_frac_generic_3xtf32_work": round(3 * f_gemm * pose_steps_per_s / (tf32_sus * 1e12), 4),
             "critic_layer2_bwd_products_per_k_step": l2p,
             "peak_tflops_tf32_sustained": tf32_sus,
             "hbm_frac_algorithmic": round(ALG_BYTES_PER_POSE_STEP * pose_steps_per_s / (pk["hbm_gbs"] * 1e9), 6),
             "useful_tflops": round((F_USEFUL if not folded else f_gemm + 2 * 24 * 17 * 33) * pose_steps_per_s / 1e12, 2),
             "flops_per_pose_step": "per-vertex formulation (SURVEY.md 8d)" if not folded else
                                    "folded formulation: two N=1224 GEMMs + critic GEMMs + the per-frame joint contraction"}

    # ---------------------------------------------------------------- regressor refit on this batch (C4 at C2's size)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for rep in range(2):
        refit.G.zero_(); refit.loss.zero_()
        torch.cuda.synchronize()
        ev[0].record()
        refit.accumulate(x6, be, gt, logical_batch=world * B)
        ev[1].record()
        if world > 1:
            dist.all_reduce(refit.G); dist.all_reduce(refit.loss)
        ev[2].record()
        refit.native.regressor_apply(refit.J, refit.mask, refit.G, refit.m, refit.v, refit.t, refit.lr)
        ev[3].record()
        torch.cuda.synchronize()
    refit_ms = {"accumulate": round(ev[0].elapsed_time(ev[1]), 3), "allreduce": round(ev[1].elapsed_time(ev[2]), 3),
                "apply": round(ev[2].elapsed_time(ev[3]), 3), "allreduce_bytes": 17 * 6890 * 4 + 4}
    refit.reset(J0)

    # ---------------------------------------------------------------- the other loss-path formulation, same run
    other = "folded" if not folded else "vertex"
    r2 = jrr.PoseRefiner(smpl, J, sd, lr=1e-2, w_joint=10000.0, w_pose=10.0, chunk=B, use_graph=True,
                         loss_path=other, steps_per_graph=U)
    r2.refine(xw.copy_(x6_0), bw.copy_(be_0), gt, iters=W)
    x2, b2 = x6_0.clone(), be_0.clone()
    t2 = timer.run(lambda: r2.refine(x2, b2, gt, iters=K))
    mp2, _ = mpjpe_of(jrr, smpl, J, x2, b2, gt, dev)
    other_path = {"loss_path": other, "value": world * B * K / (t2 * 1e-3), "unit": UNIT,
                  "ms_per_step": t2 / K, "mpjpe_after_mm": round(float(mp2), 3),
                  "gpu_launches": r2.launches_per_step * K,
                  "note": "same workload, inputs and iteration count through the other formulation of the loss path "
                          "(python bench.py --loss-path " + other + " makes it the headline)"}
    del r2
    refiner.native.set_loss_path(args.loss_path)

    # ---------------------------------------------------------------- secondary configurations (C3 / C4 / C5)
    secondary = None
    if not args.no_secondary:
        c3, c4 = run_c3(jrr, smpl, J0, sd, args, dev, rank, world, timer)
        c3["vs_device_resident_step_rate"] = round(c3["pose_steps_per_s"] / value, 4)
        c5 = run_c5(jrr, smpl, dev, pk, [1, 16, 256, 4096, 65536])
        secondary = {"c3_strong": c3, "c4_refit": c4, "c5_smpl_module": c5}
        if rank == 0:
            secondary["silhouette_term"] = run_silhouette(jrr, smpl, J0, sd, dev)

    cpu = eager_ref = None
    if rank == 0 and not args.no_cpu_baseline:
        if world == 1:
            rate, ms_cpu, cores = oracle_rate(J0, 1024, 6, 1, regressor_updates=False)
            cpu = {"value": round(rate, 1), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"1024 of the {B} frames x 6 Adam iterations (1 warm-up), {args.regressor} regressor, oracle port, "
                             f"torch {cores} threads, fp32"}
        # the honest "reference on this B200": the same oracle port in eager torch on the GPU (cuBLAS / ATen kernels)
        torch.backends.cuda.matmul.allow_tf32 = False
        rate_g, ms_g, _ = oracle_rate(J0, B, 5, 2, device=str(dev), regressor_updates=False)
        eager_ref = {"value": round(rate_g, 1), "unit": UNIT, "ms_per_step": round(ms_g, 3), "kind": "port on cuda (eager torch, fp32)",
                     "sample": f"{B} frames x 5 Adam iterations (2 warm-up), chunks of 1024, {args.regressor} regressor"}
        torch.cuda.empty_cache()

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": B, "regressor": args.regressor, "loss_path": args.loss_path,
                       "parallelism": f"frame-shard x{world}, no data-path collective in `value` (the refit all-reduce is in e2e and secondary.c3_strong)",
                       "l2": ("one folded step touches ~190 MB per GPU (activations 120 MB, split weights 33 MB, Q / dQ 42 MB)"
                              if folded else "one per-vertex step touches ~1.2 GB of intermediates per GPU") + ": larger than the 126 MB L2",
                       "graph": f"PoseRefiner.refine: CUDA graph of {U} step(s) replayed (value); one-step graph with the loss copied back per step (e2e)",
                       "gemm": "tcgen05 3xTF32" if args.gemm_impl == 0 else "simt"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                    "ms_per_step": e2e_ms / K, "last_iteration_loss_read_on_host": e2e_loss_last,
                    "note": f"per batch of <= 100 iterations: H2D of x6/betas/gt from pinned host buffers, PoseRefiner.refine with every "
                            f"iteration's loss copied back to pinned host memory, RegressorRefit.step on the refined frames "
                            f"(accumulate, all-reduce for N > 1, Adam, operator re-fold; {n_refits} in this run), D2H of the refined x6/betas"},
            "gpu_launches": main_launches * K,
            "roofline": roofline, "kernels": kern, "whole_step": whole,
            "quality": {"mpjpe_initial_mm": round(float(mp0), 3), "mpjpe_after_mm": round(float(mpjpe), 3),
                        "pa_mpjpe_after_mm": round(float(pampjpe), 3), "iterations": K, "oracle_one_step": one_step,
                        "oracle_one_step_rel_err": None if one_step is None else max(one_step["loss_rel_err"], one_step["gradient_rel_err"])},
            "refit_ms": refit_ms,
            "other_loss_path": other_path,
            "secondary": secondary,
            "cpu_baseline": cpu,
            "gpu_eager_reference": eager_ref,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
