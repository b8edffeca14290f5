"""Per-step input pipeline on the device -- SURVEY.md 8(f) row f-4.

  ShapeNetHypHC._forward / PartNetHypHC._forward   hpcs/models/shapenet_hyp_hc.py:55-91, partnet_hyp_hc.py:72-111
  to_categorical(y, num_classes)                    hpcs/utils/data.py:24-29

The reference moves every batch to the host, rotates it there with pytorch3d, uploads it again and transposes; the
one-hot category vector is built with ``torch.eye`` on the host and copied up.  Here the host draws only the rotation
PARAMETERS, from torch's CPU generator and in the reference's order (``randn(B, 4)`` for 'so3', ``rand(B)`` for 'z'),
so a common seed gives the same rotations; one kernel turns them into matrices, rotates and writes the backbone's
``[B, 3, N]`` layout (csrc/input.cu).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

ROTATION_MODES = {None: 0, "none": 0, "matrix": 1, "so3": 2, "z": 3}


def rotation_params(batch: int, mode: Optional[str]) -> Optional[torch.Tensor]:
    """The random draws the reference's rotation consumes, on the host: ``randn(B,4)`` (pytorch3d
    ``random_rotations`` -> ``random_quaternions``) for 'so3', ``rand(B)`` for 'z' (shapenet_hyp_hc.py:64-67)."""
    if mode == "so3":
        return torch.randn((batch, 4))
    if mode == "z":
        return torch.rand(batch)
    if mode in (None, "none"):
        return None
    raise ValueError(f"unknown rotation {mode!r} (expected 'so3', 'z' or None)")


def rotate_points(points: torch.Tensor, mode: Optional[str] = "so3", params: Optional[torch.Tensor] = None,
                  return_rotation: bool = False):
    """points[B,N,3] -> [B,3,N] fp32 on the device: ``(points @ R_b)^T``, the layout ``VN_DGCNN_partseg.forward``
    takes.  ``mode``: 'so3' / 'z' (``params`` = the host draws of :func:`rotation_params`; drawn here when None),
    'matrix' (``params`` = R[B,3,3]) or None (transpose only).  Points already on the device stay there; host points
    are uploaded once, un-rotated."""
    if points.dim() != 3 or points.shape[2] != 3:
        raise ValueError(f"expected points[B,N,3], got {tuple(points.shape)}")
    if mode not in ROTATION_MODES:
        raise ValueError(f"unknown rotation {mode!r}")
    code = ROTATION_MODES[mode]
    B, N, _ = points.shape
    if params is None and code in (2, 3):
        params = rotation_params(B, mode)
    if code == 1 and (params is None or tuple(params.shape) != (B, 3, 3)):
        raise ValueError("mode 'matrix' needs params = R[B,3,3]")
    if not points.is_cuda:
        raise RuntimeError("hpcs_b200 runs on CUDA (sm_100a) only; move the batch to the device first (there is no CPU fallback)")
    dev = _lib.require_cuda(points)
    lib = _lib.load()
    pts = points.detach().to(torch.float32).contiguous()
    par = None if params is None else params.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    out = torch.empty((B, 3, N), dtype=torch.float32, device=dev)
    rot = torch.empty((B, 3, 3), dtype=torch.float32, device=dev) if return_rotation else None
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_rotate_points_f32(pts.data_ptr(), _lib.ptr(par), code, B, N, out.data_ptr(), _lib.ptr(rot),
                                              _lib.stream_ptr(dev)), "hpcs_rotate_points_f32")
    return (out, rot) if return_rotation else out


def to_categorical(y: torch.Tensor, num_classes: int) -> torch.Tensor:
    """1-hot encode: ``torch.eye(num_classes)[y]`` -> ``y.shape + (num_classes,)`` fp32, on ``y``'s device without the
    host round trip of the reference."""
    dev = _lib.require_cuda(y)
    lib = _lib.load()
    yl = y.detach().to(torch.int64).contiguous()
    out = torch.empty(tuple(yl.shape) + (int(num_classes),), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_one_hot_f32(yl.data_ptr(), yl.numel(), int(num_classes), out.data_ptr(),
                                        _lib.stream_ptr(dev)), "hpcs_one_hot_f32")
    return out


def _class_vector(targets: torch.Tensor, num_parts: int) -> torch.Tensor:
    """Per cloud, which parts occur (shapenet_hyp_hc.py:77-82: one_hot(unique(targets)).sum(0)), on the device."""
    return torch.zeros((targets.shape[0], num_parts), dtype=torch.float32, device=targets.device).scatter_(1, targets, 1.0)


def _rotate_batch(self, points: torch.Tensor, testing: bool) -> torch.Tensor:
    rot = self.test_rotation if testing else self.train_rotation
    mode = rot if rot in ("so3", "z") else None
    params = rotation_params(points.shape[0], mode)                       # host RNG, the reference's draws
    return rotate_points(points.to(self.device, non_blocking=True), mode, params)


def shapenet_forward(self, batch, testing):
    """Bound onto ``ShapeNetHypHC`` (hpcs/models/shapenet_hyp_hc.py:55-91): same return tuple; rotation, transpose and
    one-hot on the device."""
    points, label, targets = batch
    points = _rotate_batch(self, points, testing)                          # [B,3,N]
    device = points.device
    label, targets = label.long().to(device), targets.long().to(device)
    if self.class_vector:
        decode_vector = _class_vector(targets, self.num_class)
    else:
        decode_vector = to_categorical(label, self.num_categories)
    x_euclidean = self.nn_feat(points, decode_vector)
    x_poincare = self.nn_emb(x_euclidean) if self.nn_emb else None
    return points, x_euclidean, x_poincare, targets


def partnet_forward(self, batch, testing):
    """Bound onto ``PartNetHypHC`` (hpcs/models/partnet_hyp_hc.py:72-111)."""
    points, targets = batch
    points = _rotate_batch(self, points, testing)
    device = points.device
    targets = targets.long().to(device)
    if self.class_vector:
        decode_vector = _class_vector(targets, self.num_class)
    else:
        decode_vector = to_categorical(torch.zeros((points.shape[0], 1), dtype=torch.long, device=device), 1)
    x_euclidean = self.nn_feat(points, decode_vector)
    x_poincare = self.nn_emb(x_euclidean) if self.nn_emb else None
    return points, x_euclidean, x_poincare, targets
