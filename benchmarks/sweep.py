#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (the headline C2 lives in bench.py):

  C1  SMPL forward + shipped regressor on 64 poses: CPU oracle (fp32) vs CUDA path, parity
  C3  312 000 synthetic frames x 100 Adam iterations, sharded over the visible ranks
      (torchrun for N > 1), chunks of 4096 with a ragged tail
  C4  regressor refit over this rank's frames: accumulate / all-reduce / apply split
  C5  SMPL forward and forward+backward, B = 2^k (k = 0..16): latency and poses/s vs the HBM
      roofline of the module path (84 172 B/pose forward, 85 076 B/pose backward)

Each result is one JSON line on stdout (rank 0); `--only c5` selects a subset.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import jrr_b200 as jrr  # noqa: E402
from bench import load_regressor  # noqa: E402


def shipped_regressor():
    """the reference artefact through the product loader (reference tree, else its byte copy under tests/golden/)"""
    return load_regressor("shipped")


def cuda_time(fn, warm=3, reps=10):
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


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (json.load(open(p))["hbm_gbs"], "measured") if os.path.exists(p) else (6650.0, "fallback")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c3,c4,cam,loop,c5")
    ap.add_argument("--frames", type=int, default=312000)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--max-log2", type=int, default=16)
    ap.add_argument("--loss-path", default="folded", choices=["vertex", "folded"])
    args = ap.parse_args()
    only = set(args.only.split(","))
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        if rank == 0:
            os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    model = jrr.synthetic.make_smpl_model(0)
    smpl = jrr.SMPL(model_dict=model, create_transl=False).to(dev)
    J = shipped_regressor()
    torch.manual_seed(0)
    sd = jrr.Discriminator().state_dict()

    if "c1" in only and rank == 0:
        from oracle import jrr_oracle as O
        inp = jrr.synthetic.make_pose_inputs(64, 0)
        R, b = torch.from_numpy(inp["true_rotmat"]), torch.from_numpy(inp["true_betas"])
        osmpl = O.OracleSMPL(model)
        torch.set_num_threads(os.cpu_count() or 1)
        ts = []
        for _ in range(12):
            t0 = time.perf_counter()
            ref, rv = O.find_joints(osmpl, b, R[:, :1], R[:, 1:], J, mask=O.find_j_reg_mask(J), return_verts=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        cpu_ms = sorted(ts[2:])[len(ts[2:]) // 2]
        Rg, bg, Jg = R.to(dev), b.to(dev), J.to(dev)
        with torch.no_grad():
            out, verts = jrr.find_joints(smpl, bg, Rg[:, :1], Rg[:, 1:], Jg, return_verts=True)
            gpu_ms_full = cuda_time(lambda: jrr.find_joints(smpl, bg, Rg[:, :1], Rg[:, 1:], Jg, return_verts=True))
            gpu_ms_fused = cuda_time(lambda: jrr.find_joints(smpl, bg, Rg[:, :1], Rg[:, 1:], Jg))
        emit({"config": "C1", "workload": "SMPL forward + retrained_J_Regressor on 64 poses, fp32",
              "cpu_oracle_ms": round(cpu_ms, 3), "cpu_poses_per_s": round(64 / cpu_ms * 1e3, 1), "cpu_cores": os.cpu_count(),
              "gpu_ms_vertices_and_joints": round(gpu_ms_full, 4), "gpu_ms_joints_only_fused": round(gpu_ms_fused, 4),
              "gpu_poses_per_s": round(64 / gpu_ms_fused * 1e3, 1),
              "vertices_rel_err": ((verts.cpu() - rv).abs().max() / rv.abs().max()).item(),
              "joints_rel_err": ((out.cpu() - ref).abs().max() / ref.abs().max()).item()})

    if "c3" in only or "c4" in only:
        lo, hi = jrr.shard_range(args.frames, rank, world)
        n = hi - lo
        inp = jrr.synthetic.make_pose_inputs(n, 1000 + rank)
        R = torch.from_numpy(inp["true_rotmat"]).to(dev)
        tb = torch.from_numpy(inp["true_betas"]).to(dev)
        gt = torch.empty(n, 17, 3, device=dev)
        with torch.no_grad():
            for c0 in range(0, n, 16384):
                c1 = min(n, c0 + 16384)
                pred = jrr.find_joints(smpl, tb[c0:c1], R[c0:c1, :1], R[c0:c1, 1:], J.to(dev))
                gt[c0:c1] = 1000 * jrr.move_pelvis(pred)
        gt += torch.from_numpy(inp["gt_noise"]).to(dev)
        x6 = torch.from_numpy(inp["x6"]).to(dev).contiguous()
        be = torch.from_numpy(inp["betas"]).to(dev).contiguous()
        del R, tb
        refiner = jrr.PoseRefiner(smpl, J, sd, chunk=4096, loss_path=args.loss_path)

        def mpjpe():
            tot, cnt = 0.0, 0
            with torch.no_grad():
                for c0 in range(0, n, 16384):
                    c1 = min(n, c0 + 16384)
                    Rg = jrr.rot6d_to_rotmat(x6[c0:c1].reshape(-1, 6)).view(-1, 24, 3, 3)
                    p = jrr.find_joints(smpl, be[c0:c1], Rg[:, :1], Rg[:, 1:], J.to(dev))
                    m, _ = jrr.evaluate(p, gt[c0:c1])
                    tot += float(m) * (c1 - c0)
                    cnt += c1 - c0
            return tot / cnt
        if "c3" in only:
            refiner.refine(x6[:4096].clone(), be[:4096].clone(), gt[:4096], iters=3)   # warm-up / graph capture
            mp0 = mpjpe()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            refiner.refine(x6, be, gt, iters=args.iters)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mp1 = mpjpe()
            emit({"config": "C3", "workload": f"{args.frames} frames x {args.iters} Adam iterations, chunk 4096 (ragged tail), "
                  f"frame-sharded over {world} GPU(s), no collective", "loss_path": args.loss_path, "n_gpus": world, "seconds": round(t.item() / 1e3, 3),
                  "pose_steps_per_s": round(args.frames * args.iters / (t.item() / 1e3), 1),
                  "mpjpe_before_mm_rank0": round(mp0, 3), "mpjpe_after_mm_rank0": round(mp1, 3)})
        if "c4" in only:
            refit = jrr.RegressorRefit(smpl, J, lr=1e-2, chunk=4096)
            J0 = refit.J_regressor.clone()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            res = []
            for step in range(3):
                refit.G.zero_(); refit.loss.zero_()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                ev[0].record()
                refit.accumulate(x6, be, gt, logical_batch=args.frames)
                ev[1].record()
                if world > 1:
                    dist.all_reduce(refit.G); dist.all_reduce(refit.loss)
                ev[2].record()
                refit.native.regressor_apply(refit.J, refit.mask, refit.G, refit.m, refit.v, refit.t, refit.lr)
                ev[3].record()
                torch.cuda.synchronize()
                res.append({"accumulate_ms": round(ev[0].elapsed_time(ev[1]), 3),
                            "allreduce_ms": round(ev[1].elapsed_time(ev[2]), 3),
                            "apply_ms": round(ev[2].elapsed_time(ev[3]), 3), "loss": refit.loss.item()})
            emit({"config": "C4", "workload": f"regressor refit (17x6890) over {args.frames} frames on {world} GPU(s); "
                  "NCCL all-reduce of 468 524 B", "n_gpus": world, "steps": res,
                  "max_abs_dJ_after_3_steps": (refit.J_regressor - J0).abs().max().item(),
                  "zero_entries_unchanged": bool(torch.equal(refit.J_regressor[J0 <= 0], J0[J0 <= 0]))})

    if "cam" in only and rank == 0:
        # widening row 8f-2: the 1000-iteration camera fit of optimize.py:187-199 on 4096 frames
        B = 4096
        inp = jrr.synthetic.make_pose_inputs(B, 11)
        x6 = torch.from_numpy(inp["x6"]).to(dev).contiguous()
        be = torch.from_numpy(inp["betas"]).to(dev).contiguous()
        gt2d = 112 + 40 * torch.randn(B, 17, 2, device=dev)
        refiner = jrr.PoseRefiner(smpl, J, sd, chunk=B)
        cam0 = torch.tensor([0.0, 0.0, 40.0], device=dev).repeat(B, 1)
        ms = cuda_time(lambda: refiner.native.camera_fit(x6, be, gt2d, cam0.clone(), 1000, 1e-2), warm=2, reps=5)
        emit({"config": "camera_fit", "workload": "4096 frames x 1000 Adam iterations on the camera translation "
              "(optimize.py:187-199); one body-model forward, then a per-frame kernel", "ms": round(ms, 3),
              "frame_iterations_per_s": round(B * 1000 / ms * 1e3)})

    if "loop" in only and rank == 0:
        # the whole per-batch loop of optimize.py:150-312 (minus SPIN inference and the silhouette term) on one
        # 4096-frame batch: camera fit 1000 it, refinement 100 it with every in-scope term, critic + shape-critic
        # training step, regressor refit step
        B = 4096
        inp = jrr.synthetic.make_pose_inputs(B, 13)
        x6 = torch.from_numpy(inp["x6"]).to(dev).contiguous()
        be = torch.from_numpy(inp["betas"]).to(dev).contiguous()
        gt = 100 * torch.randn(B, 17, 3, device=dev)
        gt2d = 112 + 40 * torch.randn(B, 17, 2, device=dev)
        torch.manual_seed(1)
        ssd = jrr.Shape_Discriminator().state_dict()
        loop = jrr.RefinementLoop(smpl, J, sd, ssd, loss_path=args.loss_path)
        cam0 = torch.tensor([0.0, 0.0, 40.0], device=dev).repeat(B, 1)
        batch = {"orient": x6[:, :1], "pose": x6[:, 1:], "betas": be, "gt_j3d": gt, "gt_j2d": gt2d, "cam": cam0}
        loop.run_batch(batch)                                           # warm-up (module loading, graphs)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        xs, bs, cs = x6.clone(), be.clone(), cam0.clone()
        gtc = jrr.move_pelvis(gt)
        torch.cuda.synchronize()
        ev[0].record()
        loop.refiner.fit_camera(xs, bs, gt2d, cs, iters=1000, logical_batch=B)
        ev[1].record()
        loop.refiner.refine_2d(xs, bs, cs, gtc, gt2d, iters=100, w_2d=0.01, logical_batch=B)
        ev[2].record()
        loop.trainer.step(xs, x6, bs, be, logical_batch=B)
        ev[3].record()
        loop.refit.step(xs, bs, gtc, logical_batch=B)
        ev[4].record()
        torch.cuda.synchronize()
        names = ["camera_fit_1000it_ms", "refine_100it_3d+2d+critic+shape_ms", "critic_training_step_ms", "regressor_refit_step_ms"]
        ms = {n: round(ev[i].elapsed_time(ev[i + 1]), 3) for i, n in enumerate(names)}
        tot = ev[0].elapsed_time(ev[4])
        emit({"config": "per_batch_loop", "workload": "optimize.py:150-312 on one 4096-frame batch (no SPIN inference, no "
              "silhouette term): camera fit, refinement (CUDA graphs of 10 iterations), critic + "
              "shape-critic training step, regressor refit", "loss_path": args.loss_path, **ms, "total_ms": round(tot, 3),
              "frames_per_s": round(B / tot * 1e3, 1)})

    if "c5" in only and rank == 0:
        hbm, src = hbm_peak()
        rows = []
        for k in range(0, args.max_log2 + 1):
            B = 1 << k
            inp = jrr.synthetic.make_pose_inputs(min(B, 4096), 7)
            rep = (B + 4095) // 4096
            R = torch.from_numpy(inp["true_rotmat"]).repeat(rep, 1, 1, 1)[:B].to(dev).contiguous()
            b = torch.from_numpy(inp["true_betas"]).repeat(rep, 1)[:B].to(dev).contiguous()
            full = R.reshape(B, 24, 9)
            nat = smpl.native()
            reps = 10 if B <= 16384 else 4
            f_ms = cuda_time(lambda: nat.smpl_forward(b, full, 0, True, True), reps=reps)
            dv = torch.randn(B, 6890, 3, device=dev)
            dj = torch.randn(B, 49, 3, device=dev)
            b_ms = cuda_time(lambda: nat.smpl_backward(b, full, 0, dv, dj), reps=reps)
            rows.append({"B": B, "fwd_us": round(f_ms * 1e3, 1), "fwd_poses_per_s": round(B / f_ms * 1e3),
                         "fwd_hbm_frac": round(84172 * B / (f_ms * 1e-3) / (hbm * 1e9), 4),
                         "fwdbwd_us": round((f_ms + b_ms) * 1e3, 1), "fwdbwd_poses_per_s": round(B / (f_ms + b_ms) * 1e3),
                         "bwd_hbm_frac": round(85076 * B / (b_ms * 1e-3) / (hbm * 1e9), 4)})
            del R, b, full, dv, dj
            torch.cuda.empty_cache()
        emit({"config": "C5", "workload": "SMPL module forward (vertices+joints49) and forward+backward, rotmat inputs",
              "hbm_peak_gbs": hbm, "peak_source": src, "rows": rows})

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
