#!/usr/bin/env python
"""Multi-GPU equality check (run under torchrun, NCCL): the per-batch loop of optimize.py:150-312
(`RefinementLoop.run`) on W ranks, each holding a shard of every global batch, must leave the same
regressor, critic and shape-critic weights and the same refined frames as ONE rank processing the
whole batches -- frames are independent, the only collectives are the gradient all-reduces of the
critic training step and of the regressor refit.  Prints one JSON line (rank 0).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 benchmarks/multi_gpu_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jrr_b200 as jrr  # noqa: E402
from conftest import shipped_regressor  # noqa: E402

N, BATCH, ITERS = 1536, 768, 5
LOSS_PATH = sys.argv[1] if len(sys.argv) > 1 else "folded"      # 'folded' (bench default) or 'vertex'


def run(model, dev, J, sd, ssd, frames, gt):
    smpl = jrr.SMPL(model_dict=model, create_transl=False).to(dev)
    loop = jrr.RefinementLoop(smpl, J, sd, ssd, refine_iters=ITERS, loss_path=LOSS_PATH)
    fr = {"orient": frames["x6"][:, :1], "pose": frames["x6"][:, 1:], "betas": frames["betas"], "gt_j3d": gt}
    hist = loop.run(fr, batch_size=BATCH)
    torch.cuda.synchronize()
    return (loop.refit.J_regressor.clone(), loop.trainer.p.clone(), loop.trainer.ps.clone(),
            [h["x6"].clone() for h in hist], [h["critic_loss"].item() for h in hist],
            [h["refit_loss"].item() for h in hist])


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    model = jrr.synthetic.make_smpl_model(0)
    J = shipped_regressor()
    torch.manual_seed(0)
    sd = jrr.Discriminator().state_dict()
    torch.manual_seed(1)
    ssd = jrr.Shape_Discriminator().state_dict()
    inp = jrr.synthetic.make_pose_inputs(N, 7)
    frames = {k: torch.from_numpy(v) for k, v in inp.items()}
    # targets: joints of the "true" parameters under the shipped regressor (CUDA path), mm, + noise
    smpl0 = jrr.SMPL(model_dict=model, create_transl=False).to(dev)
    nat = smpl0.native()
    nat.set_regressor(J.to(dev))
    with torch.no_grad():
        R = frames["true_rotmat"].to(dev).reshape(N, 24, 9)
        gt = 1000 * jrr.move_pelvis(nat.find_joints(frames["true_betas"].to(dev), R, jrr.native.POSE_ROTMAT)).cpu()
        gt = gt + frames["gt_noise"]
    # 1) whole batches on this rank alone (no process group yet -> no all-reduce)
    J1, p1, ps1, x1, lc1, lr1 = run(model, dev, J, sd, ssd, frames, gt)
    # 2) the same batches sharded over all ranks
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Jw, pw, psw, xw, lcw, lrw = run(model, dev, J, sd, ssd, frames, gt)
    dx = 0.0
    for bi, lo in enumerate(range(0, N, BATCH)):
        a, b = jrr.shard_range(min(N, lo + BATCH) - lo, rank, world)
        dx = max(dx, (xw[bi] - x1[bi][a:b]).abs().max().item())
    d = torch.tensor([(Jw - J1).abs().max().item(), (pw - p1).abs().max().item(), (pw - p1).abs().mean().item(),
                      (psw - ps1).abs().max().item(), dx], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
    dJ, dp_max, dp_mean, dps, dx = d.tolist()
    ok = dJ < 1e-5 and dp_mean < 1e-5 and dp_max < 4.1e-3 and dps < 1e-4 and dx < 2e-4
    ok = ok and all(abs(a - b) / abs(b) < 1e-4 for a, b in zip(lcw + lrw, lc1 + lr1))
    if rank == 0:
        os.write(real_stdout, (json.dumps({
            "check": "multi_gpu_equality", "loss_path": LOSS_PATH, "world": world, "frames": N, "global_batch": BATCH, "refine_iters": ITERS,
            "max_abs_diff": {"J_regressor": dJ, "critic_params_max": dp_max, "critic_params_mean": dp_mean,
                             "shape_critic_params": dps, "refined_x6": dx},
            "critic_loss": {"sharded": lcw, "single": lc1}, "refit_loss": {"sharded": lrw, "single": lr1},
            "pass": bool(ok)}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
