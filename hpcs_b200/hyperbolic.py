"""Poincare-ball ops behind the reference's signatures.

  hyp_lca(a, b, return_coord)    hpcs/distances/lca.py:37-52
  ExpMap / expmap0               hpcs/nn/hyperbolic/hyp_embed.py:6-10, hpcs/utils/poincare.py:50-54
  normalize_project (leaves)     hpcs/loss/ultrametric_loss.py:139-143 + hpcs/distances/poincare.py:61-68
"""
from __future__ import annotations

import torch

from . import _lib


def _rows(t: torch.Tensor):
    D = t.shape[-1]
    return t.reshape(-1, D), D


class _HypLca(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, return_coord):
        dev = a.device
        lib = _lib.load()
        T, D = a.shape
        out = torch.empty((T, D if return_coord else 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_hyp_lca_fwd_f32(a.data_ptr(), b.data_ptr(), T, D, int(return_coord), out.data_ptr(),
                                                _lib.stream_ptr(dev)), "hpcs_hyp_lca_fwd_f32")
        ctx.save_for_backward(a, b)
        ctx.return_coord = return_coord
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        dev = a.device
        lib = _lib.load()
        T, D = a.shape
        gout = gout.contiguous()
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_hyp_lca_bwd_f32(gout.data_ptr(), a.data_ptr(), b.data_ptr(), T, D,
                                                int(ctx.return_coord), ga.data_ptr(), gb.data_ptr(),
                                                _lib.stream_ptr(dev)), "hpcs_hyp_lca_bwd_f32")
        return ga, gb, None


def hyp_lca(a: torch.Tensor, b: torch.Tensor, return_coord: bool = True) -> torch.Tensor:
    """Projection of the origin on the geodesic through ``a`` and ``b`` ([T,D] each): its
    coordinates ([T,D]) or, with ``return_coord=False``, its hyperbolic distance to the origin
    ([T,1]).  Autograd wrt both inputs.  The scalar chain runs in fp64 on the device."""
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError(f"hyp_lca expects a,b of equal shape [T,D]; got {tuple(a.shape)}, {tuple(b.shape)}")
    _lib.require_cuda(a, b)
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError("hyp_lca: float32 only")
    return _HypLca.apply(a.contiguous(), b.contiguous(), bool(return_coord))


class _ExpMap0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        dev = u.device
        lib = _lib.load()
        rows, D = _rows(u)
        y = torch.empty_like(u)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_expmap0_fwd_f32(u.data_ptr(), rows.shape[0], D, y.data_ptr(), _lib.stream_ptr(dev)),
                       "hpcs_expmap0_fwd_f32")
        ctx.save_for_backward(u)
        return y

    @staticmethod
    def backward(ctx, gy):
        (u,) = ctx.saved_tensors
        dev = u.device
        lib = _lib.load()
        rows, D = _rows(u)
        gy = gy.contiguous()
        gu = torch.empty_like(u)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_expmap0_bwd_f32(gy.data_ptr(), u.data_ptr(), rows.shape[0], D, gu.data_ptr(),
                                                _lib.stream_ptr(dev)), "hpcs_expmap0_bwd_f32")
        return gu


def expmap0(u: torch.Tensor) -> torch.Tensor:
    """``expmap_1(u, 0)``: ``tanh(min(|u|,15)) u / max(|u|,1e-15)`` over the last dim."""
    _lib.require_cuda(u)
    if u.dtype != torch.float32:
        raise TypeError("expmap0: float32 only")
    return _ExpMap0.apply(u.contiguous())


class ExpMap(torch.nn.Module):
    """Drop-in for ``hpcs.nn.hyperbolic.ExpMap`` (hyp_embed.py:6-10)."""

    def forward(self, x):
        return expmap0(x)


def normalize_project(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """Leaves for the decoder: ``project(F.normalize(x) * clamp(scale, 1e-4, 1))`` in one kernel
    (no autograd; the reference detaches here, base_hyp_hc.py:84)."""
    dev = _lib.require_cuda(x, scale)
    lib = _lib.load()
    xc = x.detach().contiguous().float()
    rows, D = _rows(xc)
    sc = scale.detach().reshape(-1)[:1].contiguous().float()
    out = torch.empty_like(xc)
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_leaves_f32(xc.data_ptr(), rows.shape[0], D, sc.data_ptr(), out.data_ptr(),
                                       _lib.stream_ptr(dev)), "hpcs_leaves_f32")
    return out
