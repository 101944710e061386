#!/usr/bin/env python
"""Diagnostic: where does the one-step gradient of the bench configuration (B=4096, dense regressor) differ from the
fp64 oracle?  Separates the joint term from the critic term and lists the worst entries."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import jrr_b200 as jrr  # noqa: E402
from oracle import jrr_oracle as O  # noqa: E402

DEV = "cuda:0"
n, chunk = 4096, 512
model = jrr.synthetic.make_smpl_model(0)
o32, o64 = O.OracleSMPL(model), O.OracleSMPL(model, torch.float64)
J = torch.from_numpy(jrr.synthetic.make_dense_regressor(0))
sd = O.make_critic_state_dict(0)
sd64 = {k: v.double() for k, v in sd.items()}
inp = jrr.synthetic.make_pose_inputs(n, 0)
t = {k: torch.from_numpy(v) for k, v in inp.items()}
gt = torch.cat([O.make_gt(o32, J, t["true_rotmat"][lo:lo + chunk], t["true_betas"][lo:lo + chunk], t["gt_noise"][lo:lo + chunk])
                for lo in range(0, n, chunk)])
smpl = jrr.SMPL(model_dict=model, create_transl=False).to(DEV)
torch.set_num_threads(os.cpu_count())


def oracle_grad(w_joint, w_pose):
    gs = []
    for lo in range(0, n, chunk):
        x = t["x6"][lo:lo + chunk].double().requires_grad_(True)
        b = t["betas"][lo:lo + chunk].double().requires_grad_(True)
        tot, _, _, _ = O.refine_loss(o64, J.double(), sd64 if w_pose else None, x, b, gt[lo:lo + chunk].double(), w_joint=w_joint,
                                     w_pose=w_pose, logical_batch=n)
        tot.backward()
        gs.append(torch.cat([x.grad.reshape(-1, 144), b.grad], 1))
    return torch.cat(gs)


for wj, wp in ((10000.0, 10.0), (10000.0, 0.0), (0.0, 10.0)):
    g = oracle_grad(wj, wp)
    for path in ("folded", "vertex"):
        ref = jrr.PoseRefiner(smpl, J, sd if wp else None, w_joint=wj, w_pose=wp, use_graph=False, loss_path=path, chunk=n)
        st = ref._buffers(n)
        st["x6"].copy_(t["x6"]); st["betas"].copy_(t["betas"]); st["gt"].copy_(gt)
        ref._run_chunk(st, 1, n)
        torch.cuda.synchronize()
        m = st["m"].cpu().double() * 10
        d = (m - g).abs()
        gm = g.abs().max().item()
        top = torch.topk(d.flatten(), 6)
        rows = [(int(i) // 154, int(i) % 154, f"{d.flatten()[i].item() / gm:.2e}", f"g={g.flatten()[i].item():.3e}") for i in top.indices]
        per_frame = d.max(1).values / gm
        print(f"w_joint={wj} w_pose={wp} [{path}] max|g|={gm:.3e} rel err max {d.max().item() / gm:.2e} mean {d.mean().item() / gm:.2e}; "
              f"frames with err>2e-5: {(per_frame > 2e-5).sum().item()}; worst (frame,param,err,g): {rows}", flush=True)
    smpl.native().set_loss_path("vertex")
