"""On-disk input format of the refinement loop (``scripts/data.py:28-158``): one directory per split,
``data/human3.6m/precomputed_{train,val}/`` holding
``{bboxes,betas,estimated_translation,gt_j2d,gt_j3d,intrinsics,orient,pose}.pt`` (tensors, frame-major)
and ``{images,pixel_annotations}.pkl`` (python lists).  This module reads the TENSOR fields and
applies the per-frame crop arithmetic of ``data_set.__getitem__`` / ``find_crop``
(``data.py:140-158,216-270``).  The Mask R-CNN silhouettes the silhouette term compares against (``mask_rcnn`` / ``valid``,
data.py:113-131) are produced when a decoder is available: ``imageio`` as in the reference when it is installed, else a
``mask_loader(path) -> [H,W] array`` given by the caller (the mask path is derived from the frame path exactly as
data.py:115-116).  The RGB crops (``image`` / ``spin_image``) feed only the SPIN network (out of scope) and are not produced.

Host-side only (plain torch on CPU tensors): the product's device work starts at
``RefinementLoop.run_batch``.
"""
from __future__ import annotations

import os
import pickle

import torch
from torch.utils.data import Dataset

TENSOR_FILES = ("bboxes", "betas", "estimated_translation", "gt_j2d", "gt_j3d", "intrinsics", "orient", "pose")
LIST_FILES = ("images", "pixel_annotations")


def crop_window(bboxes: torch.Tensor):
    """``find_crop`` (data.py:216-247) without the image warp: bbox rows are [min_y, min_x, max_y, max_x]
    in the 1000-pixel frame.  Returns (min_x, min_y, scale, average_x, average_y): the crop's top-left
    corner in pixels, its half-size in units of 500 pixels and its centre in pixels."""
    min_x, max_x = (bboxes[:, 1] - 500) / 500, (bboxes[:, 3] - 500) / 500
    min_y, max_y = (bboxes[:, 0] - 500) / 500, (bboxes[:, 2] - 500) / 500
    ax, ay = (min_x + max_x) / 2, (min_y + max_y) / 2
    sx, sy = max_x - min_x, max_y - min_y
    scale = torch.where(sx > sy, sx, sy) / 2
    return (ax - scale) * 500 + 500, (ay - scale) * 500 + 500, scale, ax * 500 + 500, ay * 500 + 500


def crop_intrinsics(intrinsics, height, width, crop_ci, crop_cj):
    """data.py:385-410: principal point after cropping a (height x width) window centred at (ci, cj)."""
    out = intrinsics.clone()
    out[:, 0, 2] = intrinsics[:, 0, 2] + (width - 1) / 2 - crop_cj
    out[:, 1, 2] = intrinsics[:, 1, 2] + (height - 1) / 2 - crop_ci
    return out


def resize_intrinsics(intrinsics, height, width, scale):
    """data.py:413-448: intrinsics of the window resized by ``scale``."""
    out = intrinsics.clone()
    dx = intrinsics[:, 0, 2] - (width - 1) / 2
    dy = intrinsics[:, 1, 2] - (height - 1) / 2
    out[:, 0, 2] = (scale * width - 1) / 2 + scale * dx
    out[:, 1, 2] = (scale * height - 1) / 2 + scale * dy
    out[:, 0, 0] = scale * intrinsics[:, 0, 0]
    out[:, 1, 1] = scale * intrinsics[:, 1, 1]
    return out


def reposition_j2d(gt_j2d: torch.Tensor, bboxes: torch.Tensor) -> torch.Tensor:
    """data.py:134-138: 2-D joints moved into the crop and scaled to the 224-pixel render."""
    min_x, min_y, scale, _, _ = crop_window(bboxes)
    out = gt_j2d.clone()
    out[..., 0] -= min_x[:, None]
    out[..., 1] -= min_y[:, None]
    out /= scale[:, None, None]
    out /= 1000 / 224
    return out


def cropped_intrinsics(intrinsics: torch.Tensor, bboxes: torch.Tensor, img_size: int = 256) -> torch.Tensor:
    """The ``intrinsics`` field of a batch (data.py:123-124,264-268; img_size 256 is find_crop's default)."""
    _, _, scale, ax, ay = crop_window(bboxes)
    out = crop_intrinsics(intrinsics, 1000 * scale, 1000 * scale, ay, ax)
    return resize_intrinsics(out, 1000 * scale, 1000 * scale, img_size / (scale * 1000))


class data_set(Dataset):
    """``scripts.data.data_set`` for the tensor fields.  ``set`` = "train" reads
    ``<root>/precomputed_train/``, anything else ``<root>/precomputed_val/`` (data.py:31-35)."""

    def __init__(self, set, root="data/human3.6m", mask_loader=None):
        self.mask_loader = mask_loader
        if mask_loader is None:
            try:
                import imageio            # the reference's decoder (data.py:6,118-119); not part of this image
                self.mask_loader = imageio.imread
            except ImportError:
                pass
        self.location = os.path.join(root, "precomputed_train" if set == "train" else "precomputed_val")
        for name in TENSOR_FILES:
            path = os.path.join(self.location, name + ".pt")
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path}: the precomputed split is incomplete (expected {TENSOR_FILES})")
            setattr(self, name, torch.load(path, map_location="cpu", weights_only=True).detach())
        n = self.gt_j3d.shape[0]
        for name in TENSOR_FILES:
            if getattr(self, name).shape[0] != n:
                raise ValueError(f"{name}.pt has {getattr(self, name).shape[0]} frames, gt_j3d.pt has {n}")
        for name in LIST_FILES:                      # optional here: only names / annotations, no pixels
            path = os.path.join(self.location, name + ".pkl")
            setattr(self, name, pickle.load(open(path, "rb")) if os.path.exists(path) else None)
        self.inc_gt = torch.ones(n, dtype=torch.bool)             # data.py:71-75 with a single location

    def __len__(self):
        return self.gt_j3d.shape[0]

    def batch(self, index) -> dict:
        """Fields of data.py:140-158 for a tensor / list / slice of frame indices (vectorised __getitem__)."""
        if isinstance(index, slice):
            index = torch.arange(*index.indices(len(self)))
        idx = torch.as_tensor(index, dtype=torch.long).reshape(-1)
        bb = self.bboxes[idx].float()
        return {
            "bboxes": self.bboxes[idx], "betas": self.betas[idx].float(), "cam": self.estimated_translation[idx].float(),
            "gt_j2d": reposition_j2d(self.gt_j2d[idx].float(), bb), "gt_j3d": self.gt_j3d[idx].float(),
            "intrinsics": cropped_intrinsics(self.intrinsics[idx].float(), bb),
            "orient": self.orient[idx].float(), "pose": self.pose[idx].float(), "inc_gt": self.inc_gt[idx],
        }

    def masks(self, index):
        """``mask_rcnn`` [n,1,H,W] in [0,1] and ``valid`` [n] of data.py:113-131 for the given frames: the mask file is the
        frame's path with ``imageSequence`` replaced by ``maskSequence``; ``valid`` is read BEFORE the top-left 2x2 pixels
        are cleared, like the reference."""
        if self.mask_loader is None or self.images is None:
            raise RuntimeError("no mask decoder: install imageio or pass mask_loader=, and provide images.pkl")
        idx = torch.as_tensor(index, dtype=torch.long).reshape(-1).tolist()
        out = []
        for i in idx:
            head, tail = self.images[i].split("imageSequence")[:2]
            m = torch.as_tensor(self.mask_loader(f"{head}maskSequence{tail}"))
            out.append(m.to(torch.uint8).float().unsqueeze(0) / 255.0)
        mask = torch.stack(out)
        valid = mask[:, 0, 0, 0] != 0
        mask[:, :, :2, :2] = 0
        return mask, valid

    def __getitem__(self, index):
        return {k: v[0] for k, v in self.batch([int(index)]).items()}

    def batches(self, batch_size, shuffle=False, seed=0, drop_last=False):
        """Whole-batch iterator (what DataLoader + default collate give at optimize.py:136-150, without the
        per-item python loop)."""
        n = len(self)
        order = torch.randperm(n, generator=torch.Generator().manual_seed(seed)) if shuffle else torch.arange(n)
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            yield self.batch(idx)


def write_precomputed(location: str, frames: dict, images=None) -> None:
    """Writes a split directory in the reference's layout from frame-major tensors (keys = TENSOR_FILES);
    used for synthetic data and by the tests."""
    os.makedirs(location, exist_ok=True)
    for name in TENSOR_FILES:
        torch.save(frames[name].detach().cpu().clone(), os.path.join(location, name + ".pt"))
    n = frames["gt_j3d"].shape[0]
    pickle.dump(list(images) if images is not None else [f"frame_{i:06d}.jpg" for i in range(n)],
                open(os.path.join(location, "images.pkl"), "wb"))
    pickle.dump([None] * n, open(os.path.join(location, "pixel_annotations.pkl"), "wb"))
