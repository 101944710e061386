"""The pseudo-ground-truth refinement loop of ``scripts/optimize.py:144-312`` on the CUDA kernels, with every term of
``opt_loss`` (optimize.py:252-253; the silhouette term when the batch carries masks and a renderer is given):

    per batch (optimize.py:150-312)
      1. move_pelvis(gt_j3d)                                        (:162)
      2. camera-only Adam on the 2-D reprojection, 1000 iterations  (:187-199)   jrr_camera_fit
      3. 100 Adam iterations on [pose, orient, betas, cam]          (:201-265)   jrr_refine_step(_2d)
         loss = w_2d*loss_j2d + w_sil*silhouette + w_joint*joint + w_pose*pose_critic + w_shape*shape_critic
      4. critic / shape-critic training step, real = the initial (SPIN) estimates,
         fake = the refined ones                                    (:276-293)   jrr_critic_grad/apply
      5. regressor refit step on the refined meshes                 (:300-312)   jrr_regressor_*

Frames are independent: with ``torch.distributed`` initialised every rank runs the same loop on its
shard (``shard_range``) and only steps 4 and 5 all-reduce their gradient accumulators (NCCL).
The SPIN network that produces the initial estimates and the dataset loader are out of scope;
``run_batch`` takes the tensors ``optimize.py`` has in ``batch`` at line 185.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .refine import CriticTrainer, PoseRefiner, RegressorRefit, shard_range
from .native import POSE_ROT6D
from .utils import evaluate, move_pelvis


class RefinementLoop:
    def __init__(self, smpl, J_regressor, critic_state_dict, shape_critic_state_dict=None, mask=None,
                 lr=1e-2, disc_lr=1e-3, j_reg_lr=1e-2, refine_iters=100, cam_iters=1000,
                 w_joint=10000.0, w_pose=10.0, w_shape=10.0, w_2d=0.01, chunk=4096, loss_path=None,
                 silhouette_renderer=None, w_sil=100.0):
        self.native = smpl.native() if hasattr(smpl, "native") else smpl
        self.device = self.native.device
        self.refine_iters, self.cam_iters, self.w_2d = int(refine_iters), int(cam_iters), float(w_2d)
        self.renderer, self.w_sil = silhouette_renderer, float(w_sil)      # jrr_b200.Mesh_Renderer (optimize.py:111)
        # order matters: the refit owns the regressor, the trainer owns the critic weights; the refiner
        # reads both through the native model (its captured graphs stay valid across their updates)
        self.refit = RegressorRefit(smpl, J_regressor, mask=mask, lr=j_reg_lr, chunk=chunk)
        self.refiner = PoseRefiner(smpl, self.refit.J_regressor, critic_state_dict, mask=mask, lr=lr, w_joint=w_joint,
                                   w_pose=w_pose, chunk=chunk, shape_critic_state_dict=shape_critic_state_dict,
                                   w_shape=w_shape, loss_path=loss_path)
        self.trainer = CriticTrainer(smpl, critic_state_dict, shape_critic_state_dict, lr=disc_lr, chunk=chunk,
                                     w_shape=w_shape)
        self.has_shape = shape_critic_state_dict is not None

    @staticmethod
    def _world():
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def run_batch(self, batch: dict, global_batch: int | None = None) -> dict:
        """``batch`` holds THIS RANK's frames: 'orient' [n,1,6], 'pose' [n,23,6], 'betas' [n,10], 'gt_j3d'
        [n,17,3] (mm), optional 'gt_j2d' [n,17,2] + 'cam' [n,3], optional 'mask_rcnn' [n,1,S,S] (used when the loop has a
        silhouette renderer; needs the 2-D fields: the camera is fitted on them first).  ``global_batch`` is the frame count over
        all ranks (the divisor of every mean loss, optimize.py:128); default: n * world size."""
        dev = self.device
        rank, world = self._world()
        n = batch["pose"].shape[0]
        LB = n * world if global_batch is None else int(global_batch)
        x6 = torch.cat([batch["orient"].reshape(n, 1, 6), batch["pose"].reshape(n, 23, 6)], 1).to(dev).float().contiguous()
        betas = batch["betas"].to(dev).float().contiguous()
        gt = move_pelvis(batch["gt_j3d"].to(dev).float()).contiguous()
        x6_init, betas_init = x6.clone(), betas.clone()          # spin_pred_pose / spin_pred_betas
        out = {}
        use_2d = "gt_j2d" in batch and batch["gt_j2d"] is not None
        if n == 0:
            # a rank whose shard of a short last batch is empty: no local kernels, but it still takes part in the
            # all-reduces of the critic and regressor steps (zero gradient, zero loss) so the others do not block
            out["refine_loss"] = torch.zeros(5, device=dev)
            if use_2d:
                out["cam"] = batch["cam"].to(dev).float().reshape(0, 3)
        elif use_2d:
            gt2d = batch["gt_j2d"].to(dev).float().contiguous()
            cam = batch["cam"].to(dev).float().contiguous().clone()
            self.refiner.fit_camera(x6, betas, gt2d, cam, iters=self.cam_iters, logical_batch=LB)
            if self.renderer is not None and batch.get("mask_rcnn") is not None:
                loss, out["silhouette_loss"] = self.refiner.refine_silhouette(
                    x6, betas, cam, gt, gt2d, batch["mask_rcnn"], self.renderer, iters=self.refine_iters, w_2d=self.w_2d,
                    w_sil=self.w_sil, logical_batch=LB)
            else:
                loss = self.refiner.refine_2d(x6, betas, cam, gt, gt2d, iters=self.refine_iters, w_2d=self.w_2d,
                                              logical_batch=LB)
            out["cam"] = cam
        else:
            loss = self.refiner.refine(x6, betas, gt, iters=self.refine_iters, logical_batch=LB)
        if n > 0:
            out["refine_loss"] = loss.clone()
        lp, ls = self.trainer.step(x6, x6_init, betas if self.has_shape else None,
                                   betas_init if self.has_shape else None, logical_batch=LB)
        out["critic_loss"], out["shape_critic_loss"] = lp.clone(), (None if ls is None else ls.clone())
        out["refit_loss"] = self.refit.step(x6, betas, gt, logical_batch=LB).clone()
        out["x6"], out["betas"] = x6, betas
        return out

    def run(self, frames: dict, batch_size: int = 4096) -> list:
        """Whole-set driver: every global batch of ``batch_size`` frames is split over the ranks."""
        rank, world = self._world()
        N = frames["pose"].shape[0]
        hist = []
        for lo in range(0, N, batch_size):
            hi = min(N, lo + batch_size)
            a, b = shard_range(hi - lo, rank, world)
            sl = slice(lo + a, lo + b)
            hist.append(self.run_batch({k: (v[sl] if v is not None else None) for k, v in frames.items()},
                                       global_batch=hi - lo))
        return hist

    def evaluate(self, x6, betas, gt_j3d):
        """MPJPE / PA-MPJPE (mm) of the current regressor on these frames (utils.evaluate)."""
        pred = self.native.find_joints(betas.to(self.device), x6.to(self.device).reshape(-1, 24, 6), POSE_ROT6D)
        return evaluate(pred, gt_j3d.to(self.device).float())
