"""The pseudo-ground-truth refinement loop of ``scripts/optimize.py`` on the CUDA path.

``PoseRefiner.refine`` is the 100-iteration inner loop (optimize.py:201-202,220-265,
in-scope loss ``w_joint*MSE(joints) + w_pose*MSE(critic, 1)``): every iteration is one
``jrr_refine_step`` (SMPL forward + 17x6890 regressor + loss + analytic backward + Adam),
captured once into a CUDA graph and replayed.  ``RegressorRefit.step`` is the regressor
update that follows each batch (optimize.py:300-312); with ``torch.distributed`` initialised
its 17x6890 gradient accumulator is all-reduced (NCCL) so every rank applies the same step.
Frames are independent, so multi-GPU refinement is plain sharding with no communication.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .native import NativeModel


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous frame range [lo, hi) owned by `rank`."""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


class PoseRefiner:
    def __init__(self, smpl, J_regressor, critic_state_dict=None, mask=None, lr=1e-2,
                 w_joint=10000.0, w_pose=10.0, chunk=4096, use_graph=True, shape_critic_state_dict=None,
                 w_shape=10.0, loss_path=None, steps_per_graph=10):
        self.native: NativeModel = smpl.native() if hasattr(smpl, "native") else smpl
        self.device = self.native.device
        self.lr, self.w_joint = float(lr), float(w_joint)
        self.w_pose = float(w_pose) if critic_state_dict is not None else 0.0
        self.chunk = int(chunk)
        self.use_graph = use_graph
        # iterations captured per CUDA graph: > 1 removes the launch gap between consecutive replays (the step
        # counter of the Adam bias correction lives on the device, so a multi-step graph is exact); the loss
        # buffer then holds the last captured step's values
        self.steps_per_graph = max(1, int(steps_per_graph))
        if loss_path is not None:          # a property of the shared native model: the last refiner to set it wins
            self.native.set_loss_path(loss_path)
        self.set_regressor(J_regressor, mask)
        if critic_state_dict is not None:
            self.native.load_critic(critic_state_dict)
        # Shape_Discriminator term (optimize.py:244,249-250,253); the weight lives in the native model
        self.w_shape = float(w_shape) if shape_critic_state_dict is not None else 0.0
        self.native.load_shape_critic(shape_critic_state_dict, self.w_shape)
        self._bufs = {}      # B -> static buffers (+ graph)
        self.launches_per_step = 0

    def set_regressor(self, J_regressor, mask=None):
        # the refiner keeps its regressor and re-asserts it when somebody else (another refiner, a refit on
        # another tensor, utils.find_joints with another J) has replaced the shared model's copy meanwhile
        # (a device tensor is kept as it is, not copied: a refiner given RegressorRefit.J_regressor follows the refit's
        # in-place updates and the two never displace each other)
        self._J = J_regressor.detach().to(self.device)
        self._mask = None if mask is None else mask.detach().to(self.device)
        self._ensure_regressor()

    def _ensure_regressor(self):
        if not self.native.holds_regressor(self._J, self._mask):
            self.native.set_regressor(self._J, self._mask)     # bumps regressor_version: graphs are re-captured

    def _buffers(self, B, two_d=False):
        st = self._bufs.get((B, two_d))
        if st is None:
            dev = self.device
            st = {
                "x6": torch.zeros(B, 24, 6, device=dev), "betas": torch.zeros(B, 10, device=dev),
                "gt": torch.zeros(B, 17, 3, device=dev), "m": torch.zeros(B, 154, device=dev),
                "v": torch.zeros(B, 154, device=dev),
                "t": torch.zeros(1, dtype=torch.int32, device=dev),
                "loss": torch.zeros(5, device=dev), "graph": None, "LB": None,
            }
            if two_d:       # 2-D reprojection term: camera translation as a fourth Adam group
                st.update(cam=torch.zeros(B, 3, device=dev), cm=torch.zeros(B, 3, device=dev),
                          cv=torch.zeros(B, 3, device=dev), gt2d=torch.zeros(B, 17, 2, device=dev), w_2d=0.01)
            self._bufs[(B, two_d)] = st
        return st

    _STATE = ("x6", "betas", "m", "v", "t", "cam", "cm", "cv")      # what a step mutates (restored after a capture warm-up)

    def _step(self, st, LB, loss=None):
        loss = st["loss"] if loss is None else loss
        if "cam" in st:
            self.native.refine_step_2d(st["x6"], st["betas"], st["gt"], st["gt2d"], st["cam"], st["m"], st["v"], st["cm"],
                                       st["cv"], st["t"], self.lr, self.w_joint, self.w_pose, st["w_2d"],
                                       logical_batch=LB, loss_out=loss)
        else:
            self.native.refine_step(st["x6"], st["betas"], st["gt"], st["m"], st["v"], st["t"], self.lr,
                                    self.w_joint, self.w_pose, logical_batch=LB, loss_out=loss)

    def _run_chunk(self, st, iters, LB, loss_history=None):
        for k in ("m", "v", "t", "cm", "cv"):
            if k in st:
                st[k].zero_()
        if not self.use_graph:
            for i in range(iters):
                self._step(st, LB)
                if loss_history is not None:
                    loss_history[i].copy_(st["loss"], non_blocking=True)
            self.launches_per_step = self.native.launches
            return
        # everything a captured launch sequence bakes in: the packing / regressor version, the workspace pointer,
        # the loss path and shape-critic switch of the shared model, and the scalar arguments of the step
        nat = self.native
        nat.workspace(st["x6"].shape[0])      # grow (and bump the generation) BEFORE comparing, not inside the capture
        ver = (nat.regressor_version, nat.ws_generation, nat.config_generation, st.get("w_2d"), self.lr, self.w_joint,
               self.w_pose)
        if st["graph"] is None or st["LB"] != LB or st.get("ver") != ver:
            st["graphs"] = {}
            st["graph"], st["LB"], st["ver"] = self._capture(st, LB, 1), LB, ver
            st["graphs"][1] = st["graph"]
        # every captured iteration of the u-step graph writes its loss terms to its own row of st["loss_hist"], so a
        # caller that reads every iteration's loss (loss_history) still gets the multi-step graphs: one asynchronous copy
        # of u rows per replay instead of a one-step replay + copy per iteration
        u = self.steps_per_graph
        done = 0
        if u > 1 and iters >= u:
            if u not in st["graphs"]:
                st["graphs"][u] = self._capture(st, LB, u)
            for _ in range(iters // u):
                st["graphs"][u].replay()
                if loss_history is not None:
                    loss_history[done:done + u].copy_(st["loss_hist"][u], non_blocking=True)
                done += u
            st["loss"].copy_(st["loss_hist"][u][u - 1])
        for i in range(done, iters):
            st["graph"].replay()
            if loss_history is not None:
                loss_history[i].copy_(st["loss"], non_blocking=True)

    def _capture(self, st, LB, n_steps):
        # warm-up outside capture (module loading, attribute calls), on a side stream
        names = [k for k in self._STATE if k in st]
        keep = [st[k].clone() for k in names]
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self._step(st, LB)
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.launches_per_step = self.native.launches
        if n_steps > 1:      # one loss row per captured iteration (kept per graph length: a captured graph holds its pointers)
            st.setdefault("loss_hist", {}).setdefault(n_steps, torch.zeros(n_steps, 5, device=self.device))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n_steps):
                self._step(st, LB, None if n_steps == 1 else st["loss_hist"][n_steps][i])
        for k, v in zip(names, keep):
            st[k].copy_(v)
        return g

    def fit_camera(self, x6, betas, gt_j2d, cam, iters=1000, lr=1e-2, logical_batch=None):
        """optimize.py:187-199: camera-only Adam against the 2-D joints; `cam` [N,3] updated in place.
        One body-model forward per chunk (the joints do not depend on the camera)."""
        N = x6.shape[0]
        self._ensure_regressor()
        with torch.cuda.device(self.device):
            for lo in range(0, N, self.chunk):
                hi = min(N, lo + self.chunk)
                c = cam[lo:hi].contiguous()
                self.native.camera_fit(x6[lo:hi].reshape(-1, 24, 6), betas[lo:hi], gt_j2d[lo:hi], c, iters, lr,
                                       logical_batch=(hi - lo) if logical_batch is None else logical_batch)
                cam[lo:hi].copy_(c)
        return cam

    def refine_2d(self, x6, betas, cam, gt_mm, gt_j2d, iters=100, w_2d=0.01, logical_batch=None):
        """`refine` with the 2-D reprojection term and the camera as a fourth Adam group
        (optimize.py:201-202,231-233,252-253); x6 / betas / cam are updated in place.  Same chunking and
        CUDA-graph replay as `refine`."""
        N = x6.shape[0]
        last = None
        self._ensure_regressor()
        with torch.cuda.device(self.device):
            for lo in range(0, N, self.chunk):
                hi = min(N, lo + self.chunk)
                B = hi - lo
                st = self._buffers(B, two_d=True)
                st["x6"].copy_(x6[lo:hi].reshape(B, 24, 6), non_blocking=True)
                st["betas"].copy_(betas[lo:hi], non_blocking=True)
                st["cam"].copy_(cam[lo:hi], non_blocking=True)
                st["gt"].copy_(gt_mm[lo:hi].reshape(B, 17, 3), non_blocking=True)
                st["gt2d"].copy_(gt_j2d[lo:hi].reshape(B, 17, 2), non_blocking=True)
                st["w_2d"] = float(w_2d)
                self._run_chunk(st, iters, B if logical_batch is None else logical_batch)
                x6[lo:hi].copy_(st["x6"].view_as(x6[lo:hi]), non_blocking=True)
                betas[lo:hi].copy_(st["betas"], non_blocking=True)
                cam[lo:hi].copy_(st["cam"], non_blocking=True)
                last = st["loss"]
        return last

    def refine_silhouette(self, x6, betas, cam, gt_mm, gt_j2d, mask, renderer, iters=100, w_2d=0.01, w_sil=100.0,
                          logical_batch=None, use_graph=None):
        """`refine_2d` plus the silhouette term of optimize.py:234-236,252-253 (`silhouette*100`): every iteration renders
        the current mesh (`renderer`: a jrr_b200.Mesh_Renderer; module forward -> rasteriser), takes the gradient of
        `w_sil * MSE(render, mask)` back through the rasteriser and the body model (module backward) and hands it to the
        fused step as an external gradient, so ONE Adam step is taken on the sum of all terms like the reference's
        `opt_loss.backward(); optimizer.step()`.  mask [N,1,S,S] or [N,S,S].  Returns (step losses, silhouette loss) of the
        last iteration.  One iteration is ~40 launches across five C calls, issued eagerly: the host runs ahead of the device
        (2.2-2.6 ms of kernels per 1024 frames at 224 x 224), so there are no launch gaps to remove -- measured: a replayed
        iteration takes the same 2.64 ms and the capture costs 5 ms per call.  ``use_graph=True`` is kept for callers with
        a slow host: the first iteration runs eagerly, the second is captured into a CUDA graph (its intermediate tensors
        live in the graph's memory pool, so the external-gradient pointers stay valid), the rest are replays; results are
        bit-identical either way (tests/test_silhouette.py)."""
        from ._lib import POSE_ROT6D
        from .mesh_renderer import silhouette_mse
        N = x6.shape[0]
        S = renderer.image_size
        last, last_s = None, None
        self._ensure_regressor()
        nat = self.native
        with torch.cuda.device(self.device):
            for lo in range(0, N, self.chunk):
                hi = min(N, lo + self.chunk)
                B = hi - lo
                LB = B if logical_batch is None else logical_batch
                st = self._buffers(B, two_d=True)
                st["x6"].copy_(x6[lo:hi].reshape(B, 24, 6), non_blocking=True)
                st["betas"].copy_(betas[lo:hi], non_blocking=True)
                st["cam"].copy_(cam[lo:hi], non_blocking=True)
                st["gt"].copy_(gt_mm[lo:hi].reshape(B, 17, 3), non_blocking=True)
                st["gt2d"].copy_(gt_j2d[lo:hi].reshape(B, 17, 2), non_blocking=True)
                st["w_2d"] = float(w_2d)
                tgt = mask[lo:hi].reshape(B, S, S).to(self.device, torch.float32).contiguous()
                for k in ("m", "v", "t", "cm", "cv"):
                    st[k].zero_()
                def iteration():
                    verts, _ = nat.smpl_forward(st["betas"], st["x6"], POSE_ROT6D, True, False)
                    ls, dverts, dcam, _ = silhouette_mse(renderer, verts, st["cam"], tgt, LB, w_sil)
                    dbetas, dx6 = nat.smpl_backward(st["betas"], st["x6"], POSE_ROT6D, dverts, None)
                    nat.set_external_gradient(dx6, dbetas, dcam)
                    self._step(st, LB)
                    return ls

                graph = bool(use_graph)
                try:
                    done = 0
                    if iters > 0:
                        last_s = iteration()          # eager: warms up workspaces, attribute calls, tensor-map caches
                        done = 1
                    if graph and iters - done >= 2:
                        g = torch.cuda.CUDAGraph()
                        torch.cuda.synchronize(self.device)
                        with torch.cuda.graph(g):      # (the capture itself does not execute: replayed below)
                            sil_graph = iteration()
                        for _ in range(iters - done):
                            g.replay()
                        last_s = sil_graph.clone()
                        done = iters
                        torch.cuda.synchronize(self.device)      # the graph's pool (and the pointers handed to the library) die with `g`
                    for _ in range(iters - done):
                        last_s = iteration()
                finally:
                    nat.set_external_gradient()
                x6[lo:hi].copy_(st["x6"].view_as(x6[lo:hi]), non_blocking=True)
                betas[lo:hi].copy_(st["betas"], non_blocking=True)
                cam[lo:hi].copy_(st["cam"], non_blocking=True)
                last = st["loss"]
        return last, last_s

    def refine(self, x6, betas, gt_mm, iters=100, logical_batch=None, loss_history=None):
        """In-place refinement of x6 [N,24,6] / betas [N,10] against gt_mm [N,17,3] (mm,
        pelvis-centred).  Frames are processed in chunks of ``chunk``; each chunk is one
        reference "batch" (fresh Adam state; its size is the divisor of the mean losses unless
        ``logical_batch`` is given).  Returns the last iteration's [total, joint, pose, 2d, shape] loss
        of the last chunk (device tensor).

        The tensors may live on the device or in (pinned) host memory: every chunk is copied into the
        refiner's static device buffers, refined there and copied back, all asynchronously on the current
        stream.  ``loss_history`` [iters,5] (device or pinned host) receives every iteration's loss terms of
        the last chunk -- the per-iteration read-out of optimize.py:255-261 without a host sync (it costs the
        multi-step graphs: iterations are then replayed one at a time)."""
        N = x6.shape[0]
        x6v, bv, gv = x6.view(N, 24, 6), betas.view(N, 10), gt_mm.view(N, 17, 3)
        last = None
        self._ensure_regressor()
        with torch.cuda.device(self.device):
            for lo in range(0, N, self.chunk):
                hi = min(N, lo + self.chunk)
                B = hi - lo
                st = self._buffers(B)
                st["x6"].copy_(x6v[lo:hi], non_blocking=True)
                st["betas"].copy_(bv[lo:hi], non_blocking=True)
                st["gt"].copy_(gv[lo:hi], non_blocking=True)
                self._run_chunk(st, iters, B if logical_batch is None else logical_batch, loss_history)
                x6v[lo:hi].copy_(st["x6"], non_blocking=True)
                bv[lo:hi].copy_(st["betas"], non_blocking=True)
                last = st["loss"]
        return last


class RegressorRefit:
    """optimize.py:125-126,300-312 with the published no-op fixed (the raw regressor is the
    optimised tensor): persistent Adam(lr=j_reg_lr) on J_raw [17,6890]."""

    def __init__(self, smpl, J_regressor, mask=None, lr=1e-2, chunk=4096):
        self.native: NativeModel = smpl.native() if hasattr(smpl, "native") else smpl
        dev = self.native.device
        self.J = J_regressor.detach().to(dev).float().contiguous().clone()
        self.mask = None if mask is None else mask.detach().to(dev).float().contiguous()
        self.m = torch.zeros_like(self.J)
        self.v = torch.zeros_like(self.J)
        self.t = torch.zeros(1, dtype=torch.int32, device=dev)
        self.G = torch.zeros_like(self.J)
        self.loss = torch.zeros(1, device=dev)
        self.lr = float(lr)
        self.chunk = int(chunk)
        self.native.set_regressor(self.J, self.mask)

    def _ensure_regressor(self):
        # jrr_regressor_grad_accumulate / jrr_regressor_apply read the model's normalised copy: it must be the one
        # built from THIS raw regressor
        if not self.native.holds_regressor(self.J, self.mask):
            self.native.set_regressor(self.J, self.mask)

    def reset(self, J_regressor):
        """Start over from `J_regressor` (values copied into the tensor this refit owns, Adam state zeroed)."""
        self.J.copy_(J_regressor.detach().to(self.J.device).float())
        self.m.zero_(); self.v.zero_(); self.t.zero_()
        self.native.set_regressor(self.J, self.mask)

    def accumulate(self, x6, betas, gt_mm, logical_batch=None):
        """G += dL/dJhat over these frames (L = MSE with divisor 51*logical_batch).  The frames may live on the
        device or in (pinned) host memory; host chunks are uploaded asynchronously on the current stream."""
        self._ensure_regressor()
        N = x6.shape[0]
        LB = N if logical_batch is None else int(logical_batch)
        dev = self.native.device
        up = lambda t: t if t.is_cuda else t.to(dev, non_blocking=True)
        for lo in range(0, N, self.chunk):
            hi = min(N, lo + self.chunk)
            self.native.regressor_grad_accumulate(up(x6[lo:hi]), up(betas[lo:hi]), up(gt_mm[lo:hi]), self.G, self.loss,
                                                  logical_batch=LB)

    def step(self, x6=None, betas=None, gt_mm=None, logical_batch=None):
        """One refit step over this rank's frames (already refined).  With torch.distributed
        initialised the accumulators are summed over ranks first and `logical_batch` must be
        the GLOBAL frame count."""
        if x6 is not None:
            self.G.zero_()
            self.loss.zero_()
            self.accumulate(x6, betas, gt_mm, logical_batch)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.G, op=dist.ReduceOp.SUM)
            dist.all_reduce(self.loss, op=dist.ReduceOp.SUM)
        self._ensure_regressor()
        self.native.regressor_apply(self.J, self.mask, self.G, self.m, self.v, self.t, self.lr)
        return self.loss

    @property
    def J_regressor(self):
        return self.J


class CriticTrainer:
    """optimize.py:113-123,276-293: after a batch has been refined, the pose critic (and, when
    given, the shape critic) take one Adam step on MSE(D(refined), 0) + MSE(D(initial), 1).
    The flat parameter vectors, Adam moments and step counters live on the device and persist
    across batches; the model's packed critic copies are refreshed by every step, so the next
    ``PoseRefiner.refine`` sees the new weights (captured graphs stay valid: same buffers)."""

    def __init__(self, smpl, critic_state_dict, shape_critic_state_dict=None, lr=1e-3, chunk=4096, w_shape=10.0):
        from .native import flatten_critic_state_dict, flatten_shape_critic_state_dict
        self.native: NativeModel = smpl.native() if hasattr(smpl, "native") else smpl
        dev = self.native.device
        self.lr, self.chunk = float(lr), min(int(chunk), 16384)
        self.p = flatten_critic_state_dict(critic_state_dict).to(dev).contiguous()
        self.native.load_critic(critic_state_dict)
        self.ps = None
        if shape_critic_state_dict is not None:
            self.ps = flatten_shape_critic_state_dict(shape_critic_state_dict).to(dev).contiguous()
            self.native.load_shape_critic(shape_critic_state_dict, w_shape)
        mk = lambda p: None if p is None else {
            "m": torch.zeros_like(p), "v": torch.zeros_like(p), "G": torch.zeros_like(p),
            "t": torch.zeros(1, dtype=torch.int32, device=dev), "loss": torch.zeros(1, device=dev)}
        self.st, self.sts = mk(self.p), mk(self.ps)

    def _accumulate(self, st, fake, real, LB, shape):
        st["G"].zero_(); st["loss"].zero_()
        for x, target in ((fake, 0.0), (real, 1.0)):
            for lo in range(0, x.shape[0], self.chunk):
                self.native.critic_grad_accumulate(x[lo:lo + self.chunk], target, st["G"], st["loss"],
                                                   logical_batch=LB, shape=shape)

    def step(self, x6_refined, x6_initial, betas_refined=None, betas_initial=None, logical_batch=None):
        """One training step over this rank's frames.  With torch.distributed initialised the gradients
        are summed over ranks first and `logical_batch` must be the GLOBAL frame count.  Returns
        (pose critic loss, shape critic loss or None) as device tensors."""
        N = x6_refined.shape[0]
        LB = N if logical_batch is None else int(logical_batch)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        x6f, x6r = x6_refined.detach().reshape(N, 24, 6), x6_initial.detach().reshape(N, 24, 6)
        self._accumulate(self.st, x6f, x6r, LB, False)
        if self.ps is not None:
            self._accumulate(self.sts, betas_refined.detach(), betas_initial.detach(), LB, True)
        if multi:
            dist.all_reduce(self.st["G"]); dist.all_reduce(self.st["loss"])
            if self.ps is not None:
                dist.all_reduce(self.sts["G"]); dist.all_reduce(self.sts["loss"])
        self.native.critic_apply(self.p, self.st["G"], self.st["m"], self.st["v"], self.st["t"], self.lr)
        if self.ps is not None:
            self.native.critic_apply(self.ps, self.sts["G"], self.sts["m"], self.sts["v"], self.sts["t"], self.lr,
                                     shape=True)
        return self.st["loss"], (None if self.ps is None else self.sts["loss"])

    def state_dict(self):
        from .native import CRITIC_KEYS, CRITIC_SHAPES, unflatten_state_dict
        return unflatten_state_dict(self.p, CRITIC_KEYS, CRITIC_SHAPES)

    def shape_state_dict(self):
        from .native import SHAPE_CRITIC_KEYS, SHAPE_CRITIC_SHAPES, unflatten_state_dict
        return None if self.ps is None else unflatten_state_dict(self.ps, SHAPE_CRITIC_KEYS, SHAPE_CRITIC_SHAPES)


def load_j_regressor(path, device="cpu"):
    """``models/retrained_J_Regressor.pt`` (saved from cuda:0, requires_grad, column-major)
    loads unchanged: map_location + detach + contiguous (test.py:46-47 omits map_location)."""
    t = torch.load(path, map_location="cpu", weights_only=True)
    return t.detach().float().contiguous().to(device)


def save_j_regressor(J, path):
    """Plain ``torch.save(tensor)`` so the artefact stays loadable by test.py:46-47."""
    torch.save(J.detach().clone(), path)


def export_normalised_regressor(J, path=None, device=None):
    """The regressor in the form VIBE / MEVA consume it (test.py:206-208,255-256,272-273: ``nn.ReLU()(J)`` divided by
    its row sums, handed to ``model(image, J_regressor=...)``): returns relu(J) / rowsum, [17,6890] fp32 contiguous;
    with ``path`` it is also written with plain ``torch.save(tensor)`` (the artefact's own format, see
    ``save_j_regressor``).  ``J`` may be the raw tensor, a path to ``retrained_J_Regressor.pt`` or a ``RegressorRefit``.
    The mask of utils.py:182-187 is all ones and therefore not applied.  A row without a positive entry would divide
    by zero in the reference (NaN row); here that is an error."""
    if isinstance(J, (str, bytes)) or hasattr(J, "__fspath__"):
        J = load_j_regressor(J)
    elif hasattr(J, "J_regressor"):
        J = J.J_regressor
    Jn = torch.relu(J.detach().float())
    s = Jn.sum(dim=1, keepdim=True)
    if bool((s <= 0).any()):
        raise ValueError("export_normalised_regressor: a regressor row has no positive entry (the reference's "
                         "normalisation, utils.py:91-92 / test.py:207-208, would produce NaN)")
    Jn = (Jn / s).contiguous()
    if device is not None:
        Jn = Jn.to(device)
    if path is not None:
        torch.save(Jn.detach().cpu().clone(), path)
    return Jn
