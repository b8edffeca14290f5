"""Decoding embeddings into dendrograms -- drop-in for ``BaseSimilarityHypHC._decode_linkage``
(hpcs/models/base_hyp_hc.py:81-86) plus a batched form that replaces the per-cloud Python loop at
base_hyp_hc.py:135-137.

``method='complete'`` is what the reference ships (scipy ``linkage(method='complete',
metric='cosine')``); ``method='single'`` is the HypHC decoder named by the north star (same merge
order as single linkage over hyperbolic-LCA similarity, because all leaves share one radius).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .hyperbolic import normalize_project

METHODS = {"single": 0, "complete": 1}
_WS_FRACTION = 0.4             # share of the device memory the fp64 distance matrices of one chunk may take


def linkage_from_leaves(leaves: torch.Tensor, method: str = "complete") -> torch.Tensor:
    """leaves[B,N,D] fp32 on the GPU -> Z[B,N-1,4] fp64 on the GPU (scipy linkage format)."""
    if leaves.dim() != 3:
        raise ValueError(f"expected leaves[B,N,D], got {tuple(leaves.shape)}")
    if method not in METHODS:
        raise ValueError(f"method must be one of {sorted(METHODS)}")
    dev = _lib.require_cuda(leaves)
    lib = _lib.load()
    lv = leaves.detach().contiguous().float()
    B, N, D = lv.shape
    if N < 2:
        raise ValueError("need at least two leaves")
    Z = torch.empty((B, N - 1, 4), dtype=torch.float64, device=dev)
    per_cloud = lib.hpcs_linkage_workspace_bytes(1, N, D, METHODS[method])
    budget = int(torch.cuda.get_device_properties(dev).total_memory * _WS_FRACTION)
    chunk = max(1, min(B, budget // max(per_cloud, 1)))
    chunk = -(-B // -(-B // chunk))                     # equal chunks (every launch is one CTA chain per cloud)
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        ws = _lib.workspace(lib.hpcs_linkage_workspace_bytes(nb, N, D, METHODS[method]), dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_linkage_f64(lv[b0:b0 + nb].data_ptr(), nb, N, D, METHODS[method],
                                            Z[b0:b0 + nb].data_ptr(), ws.data_ptr(), ws.numel(),
                                            _lib.stream_ptr(dev)), "hpcs_linkage_f64")
    return Z


def decode_linkage_batch(x_poincare: torch.Tensor, scale: torch.Tensor, method: str = "complete",
                         return_leaves: bool = False):
    """x[B,N,D] -> Z[B,N-1,4] fp64 (GPU tensor): rescale to the common radius, project, link."""
    if x_poincare.dim() != 3:
        raise ValueError(f"expected x[B,N,D], got {tuple(x_poincare.shape)}")
    leaves = normalize_project(x_poincare, scale.to(x_poincare.device))
    Z = linkage_from_leaves(leaves, method)
    return (Z, leaves) if return_leaves else Z


def decode_linkage(leaves_embeddings: torch.Tensor, scale: torch.Tensor, method: str = "complete") -> np.ndarray:
    """One cloud [N,D] -> numpy Z[N-1,4] float64, the reference's return type."""
    Z = decode_linkage_batch(leaves_embeddings.unsqueeze(0), scale, method)
    return Z[0].cpu().numpy()


def fcluster_maxclust(Z: torch.Tensor, ks) -> torch.Tensor:
    """``scipy.cluster.hierarchy.fcluster(Z[b], k, criterion='maxclust')`` for every cloud ``b`` and every ``k`` in
    ``ks``, on the GPU: Z[B,N-1,4] (or [N-1,4]) fp64 as returned by :func:`decode_linkage_batch` -> int32 labels
    [B,K,N] (1-based, scipy's cluster numbering).  Replaces the per-k host calls of ``get_optimal_k``
    (hpcs/utils/scores.py:141-177, ``fcluster(linkage_matrix, k, criterion='maxclust')`` at :151) so Z never leaves
    the device."""
    single = Z.dim() == 2
    Zb = Z.unsqueeze(0) if single else Z
    if Zb.dim() != 3 or Zb.shape[2] != 4 or Zb.dtype != torch.float64:
        raise ValueError(f"expected Z[B,N-1,4] float64, got {tuple(Z.shape)} {Z.dtype}")
    dev = _lib.require_cuda(Zb)
    lib = _lib.load()
    ks = [int(k) for k in ks]
    if not ks or min(ks) < 1:
        raise ValueError("ks must be a non-empty list of k >= 1")
    B, M, _ = Zb.shape
    Zc = Zb.contiguous()
    ks_dev = torch.tensor(ks, dtype=torch.int32, device=dev)
    labels = torch.empty((B, len(ks), M + 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_fcluster_maxclust_i32(Zc.data_ptr(), B, M + 1, ks_dev.data_ptr(), len(ks), max(ks),
                                                  labels.data_ptr(), _lib.stream_ptr(dev)), "hpcs_fcluster_maxclust_i32")
    return labels[0] if single else labels


def get_optimal_k_batch(y: torch.Tensor, Z: torch.Tensor, index: str = "iou", extra: int = 4):
    """``get_optimal_k(y[b], Z[b], index)`` of the reference (hpcs/utils/scores.py:141-177; called per cloud on the host
    at base_hyp_hc.py:198) for a whole batch on the GPU: cut every dendrogram at k = 1 .. n_true+extra, score each cut
    against the ground-truth parts, keep the first best.  y[B,N] integer part labels (any ids), Z[B,N-1,4] fp64 on the
    device -> (best_pred[B,N] int32, 0-based cluster ids like the reference's ``fcluster(...) - 1``; best_k[B] int64;
    best_score[B] float64).  A cloud whose every score is 0 gets k = 0 and pred = -1 (the reference returns None)."""
    if index not in ("iou", "ri"):
        raise ValueError("index must be 'iou' (base_hyp_hc.py:198) or 'ri' (adjusted Rand index, viz.py:489)")
    if y.dim() != 2 or Z.dim() != 3 or y.shape[0] != Z.shape[0] or y.shape[1] != Z.shape[1] + 1:
        raise ValueError(f"expected y[B,N] and Z[B,N-1,4], got {tuple(y.shape)} and {tuple(Z.shape)}")
    dev = _lib.require_cuda(Z, y)
    lib = _lib.load()
    B, N = y.shape
    # remap_labels (scores.py:126-139): rank of a label among the labels present in its cloud
    yl = y.long()
    base = int(yl.min())
    span = int(yl.max()) - base + 1
    present = torch.zeros((B, span), dtype=torch.int32, device=dev).scatter_(1, yl - base, 1)
    ytrue = (present.cumsum(1) - 1).gather(1, yl - base).to(torch.int32).contiguous()
    n_true = present.sum(1).to(torch.int32).contiguous()
    t_cap = int(n_true.max())
    ks = list(range(1, t_cap + extra + 1))
    labels = fcluster_maxclust(Z, ks)                                     # [B,K,N]
    ks_dev = torch.tensor(ks, dtype=torch.int32, device=dev)
    scores = torch.empty((B, len(ks)), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_cut_scores_f64(labels.data_ptr(), ytrue.data_ptr(), n_true.data_ptr(), ks_dev.data_ptr(), B,
                                           len(ks), N, t_cap, max(ks), int(extra), 0 if index == "iou" else 1,
                                           scores.data_ptr(), _lib.stream_ptr(dev)), "hpcs_cut_scores_f64")
    best_score, best = scores.max(dim=1)                                  # first maximum, like the reference's strict '>'
    found = best_score > 0
    best_k = torch.where(found, best + 1, torch.zeros_like(best))
    pred = labels[torch.arange(B, device=dev), best] - 1
    pred = torch.where(found.view(B, 1), pred, torch.full_like(pred, -1))
    return pred, best_k, torch.where(found, best_score, torch.zeros_like(best_score))


def get_optimal_k(y, linkage_matrix, index):
    """Drop-in for ``hpcs.utils.scores.get_optimal_k`` (scores.py:141-177; called per cloud at base_hyp_hc.py:198 with
    ``targets[i].cpu()`` and a numpy ``Z``): same arguments, same return types -- ``(best_pred: np.ndarray | None, best_k:
    int, best_score: float)`` -- computed on the current CUDA device.  ``y`` and ``linkage_matrix`` may live on the host
    (they are uploaded: 8 N + 32 N bytes) or on the device.  For a whole batch use :func:`get_optimal_k_batch`."""
    dev = torch.device("cuda", torch.cuda.current_device())
    yt = torch.as_tensor(np.asarray(y.cpu()) if isinstance(y, torch.Tensor) and not y.is_cuda else y)
    Zt = torch.as_tensor(linkage_matrix, dtype=torch.float64)
    if yt.is_cuda:
        dev = yt.device
    elif Zt.is_cuda:
        dev = Zt.device
    pred, k, score = get_optimal_k_batch(yt.to(dev).reshape(1, -1), Zt.to(dev).unsqueeze(0), index=index)
    k = int(k[0])
    if k == 0:
        return None, 0, 0.0
    return pred[0].cpu().numpy(), k, float(score[0])
