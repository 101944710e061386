"""Pose critic with the reference's parameter layout (``scripts/discriminator.py:7-54``):
``conv_operations.{0,2}``, ``linears.{0..23}``, ``linear_operations.{0,2,4}`` -- a reference
``state_dict`` loads unchanged.  ``forward`` runs the CUDA kernels (fused 1x1 convs + heads,
3xTF32 tensor-core GEMMs for 768->1024->1024).

Limits of the drop-in (what differs from the reference ``nn.Module``):

* ``forward`` is INFERENCE ONLY: the input is detached and no autograd graph is built, neither to
  the input nor to the parameters.  The two gradients the reference loop takes through this
  module live in the C ABI instead: the input gradient of ``MSE(D(x), 1)`` inside
  ``jrr_refine_step`` (``PoseRefiner``), the parameter gradient of ``MSE(D(fake),0)+MSE(D(real),1)``
  in ``jrr_critic_grad_accumulate`` / ``jrr_critic_apply`` (``CriticTrainer``).
* ``bind(smpl.native())`` must be called first (the kernels need the device model); until then
  ``forward`` raises.  There is no PyTorch or CPU path.
* A bound module's ``forward`` writes ITS parameters into the native model's single critic slot
  whenever they changed, replacing whatever is there -- including weights a ``CriticTrainer`` on
  the same model has trained.  To score with the trained weights, load
  ``CriticTrainer.state_dict()`` into the module first (or call ``native.critic_forward``)."""
from __future__ import annotations

import torch
from torch import nn

from .native import NativeModel


class Discriminator(nn.Module):
    def __init__(self):
        super().__init__()
        self.num_inputs = 24
        self.conv_operations = nn.Sequential(nn.Conv2d(6, 32, 1), nn.ReLU(), nn.Conv2d(32, 32, 1), nn.ReLU())
        self.linears = nn.ModuleList([nn.Linear(32, 1) for _ in range(self.num_inputs)])
        self.linear_operations = nn.Sequential(nn.Linear(32 * self.num_inputs, 1024), nn.ReLU(),
                                               nn.Linear(1024, 1024), nn.ReLU(), nn.Linear(1024, 1))
        self._native = None
        self._loaded_key = None

    def bind(self, native: NativeModel):
        """Use `native`'s device copy of the weights (shared with the refinement kernels)."""
        self._native = native
        self._loaded_key = None
        return self

    def _sync_weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if key != self._loaded_key:
            self._native.load_critic(self.state_dict())
            self._loaded_key = key

    def forward(self, rot6d: torch.Tensor) -> torch.Tensor:
        """[B,24,6] -> sigmoid scores [B,25,1] ordered [global, joint 0..23]."""
        if self._native is None:
            raise RuntimeError("Discriminator.bind(smpl.native()) must be called first: the critic "
                               "runs on the CUDA kernels only")
        self._sync_weights()
        return self._native.critic_forward(rot6d.detach().reshape(-1, 24, 6)).unsqueeze(-1)


class Shape_Discriminator(nn.Module):
    """Shape critic with the reference's parameter layout (``scripts/discriminator.py:57-74``):
    ``shape_operations.{0,2,4}`` = Linear(10,10), Linear(10,5), Linear(5,1).  ``forward`` runs the
    CUDA kernel; the input gradient used by the refinement loop is inside ``jrr_refine_step``
    once the weights are handed to ``PoseRefiner(shape_critic_state_dict=...)``."""

    def __init__(self):
        super().__init__()
        self.shape_operations = nn.Sequential(nn.Linear(10, 10), nn.ReLU(), nn.Linear(10, 5), nn.ReLU(),
                                              nn.Linear(5, 1))
        self._native = None
        self._loaded_key = None
        self._w_shape = 10.0

    def bind(self, native: NativeModel, w_shape: float = 10.0):
        self._native, self._loaded_key, self._w_shape = native, None, float(w_shape)
        return self

    def forward(self, shapes: torch.Tensor) -> torch.Tensor:
        """[B,10] -> sigmoid scores [B,1]."""
        if self._native is None:
            raise RuntimeError("Shape_Discriminator.bind(smpl.native()) must be called first: the critic "
                               "runs on the CUDA kernels only")
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if key != self._loaded_key:
            self._native.load_shape_critic(self.state_dict(), self._w_shape)
            self._loaded_key = key
        return self._native.shape_critic_forward(shapes.detach().reshape(-1, 10)).unsqueeze(-1)
