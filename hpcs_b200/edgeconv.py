"""Fused EdgeConv layers of the VN-DGCNN backbone -- SURVEY.md 8(f) row f-1.

  x = get_graph_feature(x, k); x = convA(x); [x = convB(x);] x = mean_pool(x)
      hpcs/nn/dgcnn/vn_dgcnn_partseg.py:65-68, 70-73, 75-77
  VNLinearLeakyReLU / VNBatchNorm / mean_pool      hpcs/nn/dgcnn/utils/vn_layers.py:48-77, 112-132, 152-153
  VN_DGCNN_partseg.forward                          hpcs/nn/dgcnn/vn_dgcnn_partseg.py:59-103

:func:`edgeconv` takes the reference's own ``VNLinearLeakyReLU`` modules (anything with ``map_to_feat``, ``map_to_dir``
and ``batchnorm.bn``), reads their parameters, runs the fused CUDA layer (csrc/edgeconv.cu) and returns gradients for
every parameter; BatchNorm running statistics are updated like ``nn.BatchNorm2d`` does in training mode.  The edge tensor
``[B, 2C, 3, N, k]`` is never formed.  :func:`vn_dgcnn_partseg_forward` is what ``hpcs_b200.patch`` binds onto
``VN_DGCNN_partseg``: the three graph layers fused, the dense tail (conv6 ... conv11) left to the module's own layers.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .graph import knn

VO = 21                     # vector channels of every VN conv in the graph layers (64 // 3)
ROW = 128                   # floats per point row of the per-point maps
_BN, _WF, _WD, _WPAD, _W1 = 6 * VO, 256, 256 + VO * 24, 24, 256 + 2 * VO * 24


def _coef_buffer(dev) -> torch.Tensor:
    return torch.zeros(_lib.load().hpcs_edgeconv_coef_floats(), dtype=torch.float32, device=dev)


def _launch_fwd(lib, UU, VV, xd, idx, B, N, k, stages, coef, mode, stats=None, out=None, ysum=None, yrsum=None):
    dev = idx.device
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_edgeconv_fwd_f32(_lib.ptr(UU), _lib.ptr(VV), idx.data_ptr(), B, N, k, stages, coef.data_ptr(), _lib.ptr(xd), mode,
                                             _lib.ptr(stats), _lib.ptr(out), _lib.ptr(ysum), _lib.ptr(yrsum),
                                             _lib.stream_ptr(dev)), "hpcs_edgeconv_fwd_f32")


class _EdgeConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, wf1, wd1, g1, b1, wf2, wd2, g2, b2, bufs, training, momentum, eps):
        lib = _lib.load()
        dev = x.device
        B, C, _, N = x.shape
        k = idx.shape[2]
        stages = 2 if wf2 is not None else 1
        M = B * N * k
        with torch.no_grad():
            W4 = torch.stack([wf1[:, :C], wd1[:, :C], wf1[:, C:] - wf1[:, :C], wd1[:, C:] - wd1[:, :C]]).contiguous().float()
            coef = _coef_buffer(dev)
            if C == 1:                                               # coordinates layer: first conv evaluated per edge from x itself
                UU = VV = None
                xd = x
                coef[_W1:_W1 + 4 * VO] = torch.cat([wf1[:, 0], wd1[:, 0], wf1[:, 1], wd1[:, 1]]).float()
            else:
                xd = None
                UU = torch.empty((B * N, ROW), dtype=torch.float32, device=dev)
                VV = torch.empty((B * N, ROW), dtype=torch.float32, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(lib.hpcs_vn_point_linear_f32(x.data_ptr(), W4.data_ptr(), B, C, N, UU.data_ptr(), VV.data_ptr(),
                                                            _lib.stream_ptr(dev)), "hpcs_vn_point_linear_f32")
            if stages == 2:
                coef[_WF:_WF + VO * _WPAD].view(VO, _WPAD)[:, :VO] = wf2
                coef[_WD:_WD + VO * _WPAD].view(VO, _WPAD)[:, :VO] = wd2
            for s, (gamma, beta) in enumerate(((g1, b1), (g2, b2))[:stages]):
                rm, rv, nbt = bufs[3 * s:3 * s + 3]
                stats = None
                if training:                                         # batch statistics of the norms: one light pass per stage
                    stats = torch.zeros((VO, 2), dtype=torch.float64, device=dev)
                    _launch_fwd(lib, UU, VV, xd, idx, B, N, k, stages, coef, s, stats=stats)
                    if nbt is not None:
                        nbt.add_(1)
                # statistics -> folded coefficients (and the nn.BatchNorm2d running-buffer update) in one small launch
                with torch.cuda.device(dev):
                    _lib.check(lib.hpcs_edgeconv_bn_fold_f32(_lib.ptr(stats), M, gamma.data_ptr(), beta.data_ptr(), _lib.ptr(rm), _lib.ptr(rv),
                                                             _lib.ptr(nbt), -1.0 if momentum[s] is None else float(momentum[s]), float(eps[s]),
                                                             int(training), coef.data_ptr(), s, _lib.stream_ptr(dev)), "hpcs_edgeconv_bn_fold_f32")
            out = torch.empty((B, VO, 3, N), dtype=torch.float32, device=dev)
            need = any(ctx.needs_input_grad)
            ysum = torch.empty((B * N, 3 * VO), dtype=torch.float32, device=dev) if need else None
            yrsum = torch.empty_like(ysum) if need else None
            _launch_fwd(lib, UU, VV, xd, idx, B, N, k, stages, coef, 2, out=out, ysum=ysum, yrsum=yrsum)
        ctx.save_for_backward(x, idx, UU, VV, coef, ysum, yrsum, W4)
        ctx.meta = (B, C, N, k, stages, M, bool(training))
        return out

    @staticmethod
    def backward(ctx, G):
        x, idx, UU, VV, coef, ysum, yrsum, W4 = ctx.saved_tensors
        B, C, N, k, stages, M, training = ctx.meta
        xd = x if C == 1 else None
        lib = _lib.load()
        dev = x.device
        G = G.contiguous().float()
        coef = coef.clone()
        last = stages - 1
        # BatchNorm backward sums of the last stage: gy is linear in G[n]/k, the per-point coefficient sums were saved
        sums = torch.empty((VO, 2), dtype=torch.float64, device=dev)
        dg_last = torch.empty(VO, dtype=torch.float32, device=dev)
        db_last = torch.empty(VO, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_edgeconv_bn_sums_f32(G.data_ptr(), ysum.data_ptr(), yrsum.data_ptr(), B, N, k, int(training), sums.data_ptr(),
                                                     coef.data_ptr(), last, dg_last.data_ptr(), db_last.data_ptr(), _lib.stream_ptr(dev)),
                       "hpcs_edgeconv_bn_sums_f32")
        grads = {"g%d" % (last + 1): dg_last, "b%d" % (last + 1): db_last}
        gO1 = None
        with torch.cuda.device(dev):
            if stages == 2:
                gO1 = torch.empty(lib.hpcs_edgeconv_scratch_floats(B, N, k), dtype=torch.float32, device=dev)
                dW2 = torch.zeros((VO, 2, VO), dtype=torch.float32, device=dev)
                stats1 = torch.zeros((VO, 2), dtype=torch.float64, device=dev)
                _lib.check(lib.hpcs_edgeconv_bwd_stage2_f32(_lib.ptr(UU), _lib.ptr(VV), idx.data_ptr(), B, N, k, coef.data_ptr(),
                                                            _lib.ptr(xd), G.data_ptr(), gO1.data_ptr(), dW2.data_ptr(), stats1.data_ptr(),
                                                            _lib.stream_ptr(dev)), "hpcs_edgeconv_bwd_stage2_f32")
                dg1 = torch.empty(VO, dtype=torch.float32, device=dev)
                db1 = torch.empty(VO, dtype=torch.float32, device=dev)
                _lib.check(lib.hpcs_edgeconv_bn_sums_finish_f32(stats1.data_ptr(), M, int(training), coef.data_ptr(), 0, dg1.data_ptr(),
                                                                db1.data_ptr(), _lib.stream_ptr(dev)), "hpcs_edgeconv_bn_sums_finish_f32")
                grads.update(g1=dg1, b1=db1, wf2=dW2[:, 0, :], wd2=dW2[:, 1, :])
            gUU = torch.zeros((B * N, ROW), dtype=torch.float32, device=dev)
            gVV = torch.empty((B * N, ROW), dtype=torch.float32, device=dev)     # floats 63 / 127 of a row are never read
            _lib.check(lib.hpcs_edgeconv_bwd_stage1_f32(_lib.ptr(UU), _lib.ptr(VV), idx.data_ptr(), B, N, k, coef.data_ptr(),
                                                        _lib.ptr(xd), _lib.ptr(gO1), G.data_ptr(), gUU.data_ptr(), gVV.data_ptr(),
                                                        _lib.stream_ptr(dev)), "hpcs_edgeconv_bwd_stage1_f32")
        # per-point contractions (B*N points x 4 small maps): gradients wrt x and wrt the first Linear's weights, one kernel
        gx = torch.empty_like(x)
        dW4 = torch.zeros_like(W4)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_vn_point_linear_bwd_f32(gUU.data_ptr(), gVV.data_ptr(), x.data_ptr(), W4.data_ptr(), B, C, N,
                                                        gx.data_ptr(), dW4.data_ptr(), _lib.stream_ptr(dev)), "hpcs_vn_point_linear_bwd_f32")
        dwf1 = torch.cat([dW4[0] - dW4[2], dW4[2]], dim=1)
        dwd1 = torch.cat([dW4[1] - dW4[3], dW4[3]], dim=1)
        return (gx if ctx.needs_input_grad[0] else None, None, dwf1, dwd1, grads["g1"], grads["b1"], grads.get("wf2"), grads.get("wd2"),
                grads.get("g2") if stages == 2 else None, grads.get("b2") if stages == 2 else None, None, None, None, None)


def _conv_params(conv):
    bn = conv.batchnorm.bn
    if conv.map_to_dir.weight.shape[0] != conv.map_to_feat.weight.shape[0]:
        raise ValueError("fused EdgeConv needs share_nonlinearity=False (one direction per output channel)")
    if abs(float(conv.negative_slope) - 0.2) > 1e-12:
        raise ValueError("fused EdgeConv is built for negative_slope = 0.2 (the backbone's value)")
    if bn.weight is None or bn.running_mean is None:
        raise ValueError("fused EdgeConv needs an affine BatchNorm that tracks running statistics")
    return bn


def edgeconv(x: torch.Tensor, k: int, conv_a, conv_b=None, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``mean_pool(conv_b(conv_a(get_graph_feature(x, k))))`` in one fused layer: x[B,C,3,N] -> [B,21,3,N].

    ``conv_a`` / ``conv_b``: ``VNLinearLeakyReLU`` modules (2C -> 21 and 21 -> 21; ``conv_b`` None for a one-conv layer);
    train / eval mode is taken from ``conv_a.training``.  ``idx`` overrides the dynamic kNN graph, like the reference's
    ``get_graph_feature(x, k, idx)``.  Autograd wrt ``x`` and every parameter of the two modules."""
    if x.dim() != 4 or x.shape[2] != 3:
        raise ValueError(f"expected x[B,C,3,N], got {tuple(x.shape)}")
    dev = _lib.require_cuda(x, idx)
    if x.dtype != torch.float32:
        raise TypeError("edgeconv: float32 only")
    B, C, _, N = x.shape
    bn_a = _conv_params(conv_a)
    bn_b = _conv_params(conv_b) if conv_b is not None else None
    wf1, wd1 = conv_a.map_to_feat.weight, conv_a.map_to_dir.weight
    if tuple(wf1.shape) != (VO, 2 * C) or (conv_b is not None and tuple(conv_b.map_to_feat.weight.shape) != (VO, VO)):
        raise ValueError(f"fused EdgeConv handles 2C -> 21 [-> 21] channel maps; got {tuple(wf1.shape)} for C={C}")
    xc = x.contiguous()
    if idx is None:
        idx = knn(xc.detach().view(B, 3 * C, N), k)
    else:
        if tuple(idx.shape) != (B, N, k):
            raise ValueError(f"idx must be [B,N,k]={B, N, k}, got {tuple(idx.shape)}")
        trusted = getattr(idx, "_hpcs_knn_of", None) == N
        idx = idx.to(device=dev, dtype=torch.int64).contiguous()
        if not trusted:
            lo, hi = torch.aminmax(idx)
            torch._assert_async((lo >= 0) & (hi < N), "edgeconv: idx out of range [0, N)")
    bufs = [bn_a.running_mean, bn_a.running_var, bn_a.num_batches_tracked]
    bufs += [bn_b.running_mean, bn_b.running_var, bn_b.num_batches_tracked] if bn_b is not None else [None, None, None]
    mom = (bn_a.momentum, bn_b.momentum if bn_b is not None else None)
    eps = (bn_a.eps, bn_b.eps if bn_b is not None else 0.0)
    training = conv_a.training if bn_a.track_running_stats else True
    return _EdgeConv.apply(xc, idx, wf1, wd1, bn_a.weight, bn_a.bias,
                           conv_b.map_to_feat.weight if conv_b is not None else None,
                           conv_b.map_to_dir.weight if conv_b is not None else None,
                           bn_b.weight if bn_b is not None else None, bn_b.bias if bn_b is not None else None,
                           bufs, training, mom, eps)


def vn_dgcnn_partseg_forward(self, x, l):
    """Bound onto ``VN_DGCNN_partseg`` (hpcs/nn/dgcnn/vn_dgcnn_partseg.py:59-103) when its pooling is 'mean' (what
    ``train.py:68`` constructs): x[B,3,N], l[B,num_categories(,1)] -> [B,N,out_features].  The three graph layers run
    fused; everything after them is the module's own dense layers, called in the reference's order."""
    if self.pooling != "mean":                               # VNMaxPool variant: unfused layers over the native kNN + gather
        return self._hpcs_reference_forward(x, l)
    B, _, N = x.shape
    x1 = edgeconv(x.unsqueeze(1), self.k, self.conv1, self.conv2)
    x2 = edgeconv(x1, self.k, self.conv3, self.conv4)
    x3 = edgeconv(x2, self.k, self.conv5)
    x123 = torch.cat((x1, x2, x3), dim=1)                    # [B,63,3,N]
    feat = self.conv6(x123)
    feat = torch.cat((feat, feat.mean(dim=-1, keepdim=True).expand(feat.size())), 1)
    feat, frame = self.std_feature(feat)                      # invariant features + the frame z0
    x123 = torch.einsum("bijm,bjkm->bikm", x123, frame).reshape(B, -1, N)
    pooled = feat.reshape(B, -1, N).max(dim=-1, keepdim=True)[0]
    cat = self.conv7(l.view(B, -1, 1))
    glob = torch.cat((pooled, cat), dim=1).repeat(1, 1, N)
    h = torch.cat((glob, x123), dim=1)
    h = self.dp1(self.conv8(h))
    h = self.dp2(self.conv9(h))
    h = self.conv11(self.conv10(h))
    return h.transpose(1, 2)
