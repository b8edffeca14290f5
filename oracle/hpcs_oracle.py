"""CPU oracle for the HPCS hot path (kNN graph, Poincare triplet objective, linkage decode).

TEST INFRASTRUCTURE ONLY -- never imported by the product package ``hpcs_b200``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it.

Each function restates, in plain PyTorch / numpy / scipy, what the reference computes; the
docstring names the reference file:line it follows (paths relative to ``/root/reference``).  All
functions are dtype-generic: feed ``float64`` tensors to get the fp64 evaluation that SURVEY.md
(Finding 4) defines as the parity target for distances, losses and gradients.

PARITY PIN: the reference ships no golden vectors.  The oracle is pinned
instead against outputs of the reference itself, executed unmodified in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (checked by
``tests/test_oracle_golden.py``).  scipy (``scipy.cluster.hierarchy.linkage``) is the reference's
own third-party decoder (``hpcs-env.yaml:302`` pins 1.9.1; 1.18.1 in this image) and is called
directly.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MIN_NORM = 1e-15                       # hpcs/distances/poincare.py:9
BALL_EPS = {torch.float32: 4e-3, torch.float64: 1e-5}   # hpcs/distances/poincare.py:10
ARTANH_CLAMP = 1e-5                    # hpcs/utils/math.py:64
SCALE_MIN, SCALE_MAX = 1e-4, 1.0       # hpcs/loss/ultrametric_loss.py:141-143


# --------------------------------------------------------------------------------------------
# part 1: kNN graph + edge features
# --------------------------------------------------------------------------------------------
def knn_reference(x: torch.Tensor, k: int) -> torch.Tensor:
    """``knn`` as the reference evaluates it (hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:4-10).

    x[B,D,N] -> idx[B,N,k] int64: top-k of ``-|xi|^2 + 2 xi.xj - |xj|^2`` per row (self included),
    using torch's matmul/topk, i.e. with whatever accumulation order and tie order they pick.
    """
    gram = torch.matmul(x.transpose(2, 1), x)
    sq = (x * x).sum(dim=1, keepdim=True)
    neg_d2 = -sq - (-2 * gram) - sq.transpose(2, 1)
    return neg_d2.topk(k=k, dim=-1)[1]


def neg_sqdist_fp64(x: torch.Tensor) -> torch.Tensor:
    """Exact-ish ``-|xi-xj|^2`` in fp64 from fp32 inputs; used to grade near-ties (Finding 5)."""
    xd = x.double()
    diff = xd.unsqueeze(3) - xd.unsqueeze(2)          # [B,D,N,N]
    return -(diff * diff).sum(dim=1)


_KNN_LIB = None


def _knn_lib():
    global _KNN_LIB
    if _KNN_LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "libknn_canonical.so")
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build()
        lib = ctypes.CDLL(path)
        lib.knn_canonical_f32.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.knn_canonical_f32.restype = ctypes.c_int
        _KNN_LIB = lib
    return _KNN_LIB


def knn_canonical(x: torch.Tensor, k: int, return_values: bool = False):
    """Canonical-order kNN (C restatement in ``oracle/knn_canonical.c``).

    Same quantity as vn_dgcnn_util.py:5-9, evaluated in fp32 with ONE fixed operation order
    (fma chain over d ascending) and the stated tie-break (larger value first, then lower index).
    This is the bit-exact target of the CUDA kernel.
    """
    assert x.dtype == torch.float32 and x.dim() == 3
    xc = x.detach().cpu().contiguous()
    B, D, N = xc.shape
    idx = torch.empty(B, N, k, dtype=torch.int64)
    val = torch.empty(B, N, k, dtype=torch.float32)
    rc = _knn_lib().knn_canonical_f32(xc.data_ptr(), B, D, N, k, idx.data_ptr(), val.data_ptr())
    if rc != 0:
        raise RuntimeError(f"knn_canonical_f32 failed: {rc}")
    return (idx, val) if return_values else idx


def graph_feature(x: torch.Tensor, k: int = 20, idx: Optional[torch.Tensor] = None,
                  x_coord: Optional[torch.Tensor] = None, cross: bool = False,
                  knn_fn=knn_reference) -> torch.Tensor:
    """Edge features (vn_dgcnn_util.py:13-41; cross variant :44-69).

    x[B,C,3,N] -> [B,2C,3,N,k] = cat(x_j - x_i, x_i) over vector channels (+ cross(x_j, x_i) ->
    3C channels when ``cross``).  Written with plain gathers; differentiable wrt x.
    """
    B, C, _, N = x.shape
    flat = x.reshape(B, C * 3, N)
    if idx is None:
        idx = knn_fn(flat if x_coord is None else x_coord, k)
    rows = flat.transpose(1, 2).reshape(B * N, 3 * C)                   # one row per point
    base = torch.arange(B, device=idx.device).view(B, 1, 1) * N          # vn_dgcnn_util.py:25-29
    nbr = rows.index_select(0, (idx + base).reshape(-1))                # [B*N*k, 3C]
    nbr = nbr.reshape(B, N, k, C, 3)
    ctr = rows.reshape(B, N, 1, C, 3).expand(B, N, k, C, 3)
    parts = [nbr - ctr, ctr]
    if cross:
        parts.append(torch.cross(nbr, ctr, dim=-1))
    return torch.cat(parts, dim=3).permute(0, 3, 4, 1, 2).contiguous()


# --------------------------------------------------------------------------------------------
# part 2: Poincare math, hyperbolic LCA, triplet objective
# --------------------------------------------------------------------------------------------
class _Artanh(torch.autograd.Function):
    """hpcs/utils/math.py:61-74: clamp to +-(1-1e-5), forward in double, backward g/(1-x_c^2)."""

    @staticmethod
    def forward(ctx, x):
        xc = x.clamp(-1 + ARTANH_CLAMP, 1 - ARTANH_CLAMP)
        ctx.save_for_backward(xc)
        z = xc.double()
        return (0.5 * (torch.log1p(z) - torch.log1p(-z))).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        return g / (1 - xc * xc)


def artanh(x):
    return _Artanh.apply(x)


def tanh_clamped(x):
    """hpcs/utils/math.py:81-82."""
    return x.clamp(-15, 15).tanh()


def expmap0(u: torch.Tensor) -> torch.Tensor:
    """``ExpMap.forward`` = ``expmap_1(u, 0)`` (hpcs/nn/hyperbolic/hyp_embed.py:6-10,
    hpcs/utils/poincare.py:50-54,79-86).  With p = 0 the conformal factor is 2 and the Moebius
    addition ``0 (+) y`` is ``y / max(1, MIN_NORM)``."""
    nrm = u.norm(dim=-1, p=2, keepdim=True).clamp_min(MIN_NORM)
    y = tanh_clamped(2.0 * nrm / 2) * u / nrm
    p = torch.zeros_like(u)
    return mobius_add(p, y)


def mobius_add(x, y):
    """hpcs/utils/poincare.py:79-86 (same body: hpcs/distances/poincare.py:71-78)."""
    x2 = (x * x).sum(-1, keepdim=True)
    y2 = (y * y).sum(-1, keepdim=True)
    xy = (x * y).sum(-1, keepdim=True)
    top = (1 + 2 * xy + y2) * x + (1 - x2) * y
    return top / (1 + 2 * xy + x2 * y2).clamp_min(MIN_NORM)


def project(x):
    """hpcs/distances/poincare.py:61-68."""
    nrm = x.norm(dim=-1, p=2, keepdim=True).clamp_min(MIN_NORM)
    lim = 1 - BALL_EPS[x.dtype]
    return torch.where(nrm > lim, x / nrm * lim, x)


def hyp_dist_o(x):
    """hpcs/distances/poincare.py:131-136."""
    return 2 * artanh(x.norm(dim=-1, p=2, keepdim=True))


def _invert(center, x):
    """Circle inversion through the circle centred at ``center`` orthogonal to the unit sphere
    (hpcs/distances/lca.py:8-12)."""
    rad2 = (center * center).sum(-1, keepdim=True) - 1.0
    u = x - center
    return rad2 / (u * u).sum(-1, keepdim=True) * u + center


def hyp_lca(a, b, return_coord: bool = True):
    """Projection of the origin on the geodesic through a and b (hpcs/distances/lca.py:37-52,
    helpers :15-34)."""
    r = a / (a * a).sum(-1, keepdim=True)                              # :15-17
    b_inv = _invert(r, b)                                              # :44
    # reflect a across the line through the origin and b_inv (:20-29)
    dot = (a * b_inv).sum(-1, keepdim=True)
    nb = (b_inv * b_inv).sum(-1, keepdim=True).clamp_min(MIN_NORM)
    a_ref = 2 * (dot * b_inv / nb) - a
    o_ref = _invert(r, a_ref)                                          # :47
    proj = o_ref / (1.0 + torch.sqrt(1 - (o_ref * o_ref).sum(-1, keepdim=True)))   # :32-34
    return proj if return_coord else hyp_dist_o(proj)


def cosine_similarity_matrix(q, r=None):
    """``CosineSimilarity()(q, r)`` (hpcs/distances/cosine.py:4-16 on PML ``BaseDistance``:
    rows are L2-normalised with ``F.normalize`` eps 1e-12 first)."""
    qn = F.normalize(q, p=2, dim=1)
    rn = qn if r is None else F.normalize(r, p=2, dim=1)
    return 0.5 * (1 + qn @ rn.t())


def normalize_embeddings(e, scale):
    """hpcs/loss/ultrametric_loss.py:139-143."""
    return F.normalize(e, p=2, dim=1) * torch.clamp(scale, SCALE_MIN, SCALE_MAX)


def sample_triplets(labels: torch.Tensor, t_per_anchor: int, fraction: float
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Class-balanced random triplets (hpcs/miner/loss_and_miner_utils.py:7-75, ``ref_labels is
    labels`` and ``weights is None`` branch).  Consumes the global torch CPU RNG in the same
    order as the reference (one ``randint`` for positives, one for negatives, per label in
    ascending label order), so a common seed yields identical triplets."""
    lab = labels.cpu()
    counts = torch.bincount(lab)
    biggest = counts.max()
    out_a, out_p, out_n = [], [], []
    for value in torch.unique(lab):
        members = torch.nonzero(lab == value, as_tuple=True)[0]
        others = torch.nonzero(lab != value, as_tuple=True)[0]
        m = members.numel()
        if m < 2 or others.numel() < 1:
            continue
        per_anchor = int(t_per_anchor * torch.pow(biggest / m, fraction))     # :30
        total = m * per_anchor
        draw_p = torch.randint(0, m - 1, (total,))                            # :38
        anchor_pos = torch.arange(m).repeat_interleave(per_anchor)            # :40
        # the reference removes the diagonal of an m x m table; skipping slot ``anchor_pos``:
        pos_pos = draw_p + (draw_p >= anchor_pos).long()
        draw_n = torch.randint(0, others.numel(), (total,))                   # :61
        out_a.append(members[anchor_pos])
        out_p.append(members[pos_pos])
        out_n.append(others[draw_n])
    if not out_a:
        e = torch.empty(0, dtype=torch.long, device=labels.device)
        return e, e.clone(), e.clone()
    dev = labels.device
    return torch.cat(out_a).to(dev), torch.cat(out_p).to(dev), torch.cat(out_n).to(dev)


def filter_triplets(x, a, p, n, margin: float = 0.0, kind: str = "easy"):
    """``RandomTripletMarginMiner.mine`` after sampling (hpcs/miner/triplet_margin_miner.py:16-38);
    the similarity is inverted, so the margin is ``sim(a,p) - sim(a,n)``."""
    with torch.no_grad():
        mat = cosine_similarity_matrix(x)
        gap = mat[a, p] - mat[a, n]
        if kind == "easy":
            keep = gap > margin
        else:
            keep = gap <= margin
            if kind == "hard":
                keep &= gap <= 0
            elif kind == "semihard":
                keep &= gap > 0
    return a[keep], p[keep], n[keep]


def compute_hyp(x, a, p, n, scale, temperature: float):
    """``MetricHyperbolicLoss.compute_hyp`` after mining (hpcs/loss/ultrametric_loss.py:64-93):
    dense similarity matrix, three LCA distances per triplet, softmax-weighted HypHC loss plus
    the mean of the similarity matrix."""
    sim = cosine_similarity_matrix(x)                                   # :65
    w = torch.stack([sim[a, p], sim[a, n], sim[p, n]], dim=-1)          # :67-69,:84
    ea, ep, en = (normalize_embeddings(x[i], scale) for i in (a, p, n))  # :71-77
    d = torch.cat([hyp_lca(ea, ep, False), hyp_lca(ea, en, False), hyp_lca(ep, en, False)], dim=-1)
    soft = torch.softmax(d / temperature, dim=-1)                       # :86
    per_triplet = w.sum(-1) - (w * soft).sum(-1)                        # :88-89
    return per_triplet.mean() + sim.mean()                              # :91


def cosface_logits(embeddings: torch.Tensor, W: torch.Tensor, labels: torch.Tensor, margin: float, scale: float) -> torch.Tensor:
    """``MetricHyperbolicLoss.get_logits`` (hpcs/loss/ultrametric_loss.py:95-112), statement by statement, with the members of
    pytorch-metric-learning 1.6.3 ``CosFaceLoss`` it calls written out (third-party, not under /root/reference: ``get_cosine``
    = cosine of the L2-normalised rows and columns, ``get_target_mask`` = one-hot, ``modify_cosine_of_target_classes`` =
    ``c - margin``, ``scale_logits`` = ``logits * scale``)."""
    mask = F.one_hot(labels.long(), W.shape[1]).to(embeddings.dtype)                     # :99
    cosine = F.normalize(embeddings, p=2, dim=1) @ F.normalize(W, p=2, dim=0)           # :100
    cosine_of_target_classes = cosine[mask == 1]                                        # :101
    modified = cosine_of_target_classes - margin                                        # :102-104
    diff = (modified - cosine_of_target_classes).unsqueeze(1)                           # :105-107
    logits = cosine + (mask * diff)                                                     # :108
    return logits * scale                                                               # :109


# --------------------------------------------------------------------------------------------
# part 3: linkage decode
# --------------------------------------------------------------------------------------------
def decode_linkage(x: torch.Tensor, scale: torch.Tensor, method: str = "complete") -> np.ndarray:
    """``BaseSimilarityHypHC._decode_linkage`` (hpcs/models/base_hyp_hc.py:81-86): rescale to the
    common radius, project, go to CPU, scipy ``linkage(metric='cosine')``.  ``method='complete'``
    is what the reference ships; ``'single'`` is the HypHC-paper decoder named by north_star."""
    from scipy.cluster.hierarchy import linkage
    e = project(normalize_embeddings(x, scale)).detach().cpu()
    return linkage(e.numpy(), method=method, metric="cosine")


# --- step-by-step restatement of what scipy does inside ``linkage`` (small inputs only) ---------
# scipy is a compiled third-party dependency of the reference (not vendored under /root/reference);
# its published algorithms (pdist 'cosine'; Prim-style ``mst_single_linkage``; ``nn_chain`` with the
# complete-linkage update; stable sort by height; union-find ``label``) are restated here so the
# CUDA decoder has a line-by-line checker, and are themselves checked bit-for-bit against
# ``scipy.cluster.hierarchy.linkage`` in tests/test_oracle_golden.py.
def pdist_cosine_restated(X: np.ndarray) -> np.ndarray:
    """fp64 cosine distance, full symmetric matrix.  Arithmetic of the scipy build in this image
    (measured): dot products use TWO running sums (even / odd elements, separate
    multiply and add, no FMA), added at the end, odd tail element added last;
    ``1 - dot / (norm_i * norm_j)`` with ``|cos| > 1`` clipped to +-1."""
    X = np.asarray(X, dtype=np.float64)
    n, D = X.shape

    def dot2(u, v):
        even = np.float64(0.0)
        odd = np.float64(0.0)
        for q in range(0, D - 1, 2):
            even = even + u[q] * v[q]
            odd = odd + u[q + 1] * v[q + 1]
        s = even + odd
        if D % 2:
            s = s + u[D - 1] * v[D - 1]
        return s

    norms = np.array([np.sqrt(dot2(X[i], X[i])) for i in range(n)])
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            c = dot2(X[i], X[j]) / (norms[i] * norms[j])
            if abs(c) > 1.0:
                c = np.copysign(1.0, c)
            out[i, j] = out[j, i] = 1.0 - c
    return out


def _label(rows, n):
    """Union-find relabel: smaller root id first, new cluster id n+i, size in column 3."""
    parent = list(range(2 * n - 1))
    size = [1] * n + [0] * (n - 1)

    def find(v):
        root = v
        while parent[root] != root:
            root = parent[root]
        while parent[v] != root:
            parent[v], v = root, parent[v]
        return root

    Z = np.zeros((n - 1, 4))
    for i, (x, y, h) in enumerate(rows):
        rx, ry = find(int(x)), find(int(y))
        lo, hi = (rx, ry) if rx < ry else (ry, rx)
        new = n + i
        parent[rx] = parent[ry] = new
        size[new] = size[rx] + size[ry]
        Z[i] = (lo, hi, h, size[new])
    return Z


def linkage_restated(dm: np.ndarray, method: str) -> np.ndarray:
    """Dendrogram from a full fp64 distance matrix, following scipy's two code paths."""
    n = dm.shape[0]
    rows = []
    if method == "single":                       # Prim from node 0, strict '<' -> lowest index on ties
        best = np.full(n, np.inf)
        done = np.zeros(n, dtype=bool)
        x = 0
        for _ in range(n - 1):
            done[x] = True
            cur, y = np.inf, 0
            for i in range(n):
                if done[i]:
                    continue
                if best[i] > dm[x, i]:
                    best[i] = dm[x, i]
                if best[i] < cur:
                    cur, y = best[i], i
            rows.append((x, y, cur))
            x = y
    elif method == "complete":                   # nearest-neighbour chain, Lance-Williams max update
        D = dm.copy()
        size = np.ones(n, dtype=np.int64)
        chain = []
        for _ in range(n - 1):
            if not chain:
                chain.append(int(np.nonzero(size > 0)[0][0]))
            while True:
                x = chain[-1]
                if len(chain) > 1:
                    y, cur = chain[-2], D[x, chain[-2]]
                else:
                    y, cur = 0, np.inf
                for i in range(n):
                    if size[i] == 0 or i == x:
                        continue
                    if D[x, i] < cur:
                        cur, y = D[x, i], i
                if len(chain) > 1 and y == chain[-2]:
                    break
                chain.append(y)
            chain = chain[:-2]
            if x > y:
                x, y = y, x
            rows.append((x, y, cur))
            nx, ny = size[x], size[y]
            size[x], size[y] = 0, nx + ny
            for i in range(n):
                if size[i] == 0 or i == y:
                    continue
                D[i, y] = D[y, i] = max(D[i, x], D[i, y])
    else:
        raise ValueError(method)
    order = np.argsort(np.array([r[2] for r in rows]), kind="stable")
    return _label([rows[i] for i in order], n)


def fcluster_maxclust_restated(Z: np.ndarray, k: int) -> np.ndarray:
    """scipy.cluster.hierarchy.fcluster(Z, k, 'maxclust') restated (scipy/cluster/_hierarchy.pyx: cluster_maxclust_monocrit
    + cluster_monocrit) for a monotone Z, the way csrc/cut.cu evaluates it: the cut merges the fewest rows that leave
    <= k clusters (whole runs of equal heights), then the depth-first numbering restricted to the rows above the
    cutoff, subtrees filled through the dendrogram's leaf order.  Checked against scipy itself in tests/test_oracle_golden.py."""
    Z = np.asarray(Z, dtype=np.float64)
    M = Z.shape[0]
    N = M + 1
    if k >= N:                          # scipy's shortcut: every point its own cluster, numbered by point index
        return np.arange(1, N + 1, dtype=np.int32)
    left, right, size, h = Z[:, 0].astype(int), Z[:, 1].astype(int), Z[:, 3].astype(int), Z[:, 2]
    lo = np.zeros(2 * N - 1, dtype=int)
    for i in range(M - 1, -1, -1):
        lo[left[i]] = lo[N + i]
        lo[right[i]] = lo[N + i] + (1 if left[i] < N else size[left[i] - N])
    order = np.zeros(N, dtype=int)
    order[lo[:N]] = np.arange(N)
    j = N - k - 1                       # merging rows 0..j leaves k clusters
    c = 0                               # rows with height <= cutoff: [0, c)
    if j >= 0:
        c = j + 1
        while c < M and h[c] == h[j]:   # a cut cannot split a run of equal heights: fewer than k clusters then
            c += 1
    clusters = []
    if M - 1 < c:
        clusters.append(N + M - 1)
    else:
        stack = [[M - 1, 0]]
        while stack:
            row, st = stack[-1]
            l, r = left[row], right[row]
            if st == 0:
                stack[-1][1] = 1
                if l >= N:
                    if l - N >= c:
                        stack.append([l - N, 0]); continue
                    clusters.append(l)
            if stack[-1][1] == 1:
                stack[-1][1] = 2
                if r >= N:
                    if r - N >= c:
                        stack.append([r - N, 0]); continue
                    clusters.append(r)
            if l < N:
                clusters.append(l)
            if r < N:
                clusters.append(r)
            stack.pop()
    T = np.zeros(N, dtype=np.int32)
    for cid, node in enumerate(clusters, start=1):
        cnt = 1 if node < N else size[node - N]
        T[order[lo[node]:lo[node] + cnt]] = cid
    return T


def adjusted_rand_from_contingency(C: np.ndarray) -> float:
    """sklearn.metrics.adjusted_rand_score (the reference's ``ri``) from the contingency table, in Python integers like
    sklearn's pair_confusion_matrix + adjusted_rand_score."""
    C = C.astype(object)
    n = int(C.sum())
    nk, nc = C.sum(1), C.sum(0)
    sumsq = int((C * C).sum())
    tp = sumsq - n
    fp = int((C * nc[None, :]).sum()) - sumsq
    fn = int((C * nk[:, None]).sum()) - sumsq
    tn = n * n - fp - fn - sumsq
    if fn == 0 and fp == 0:
        return 1.0
    return 2.0 * (tp * tn - fn * fp) / ((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn))


def get_optimal_k_restated(y: np.ndarray, Z: np.ndarray, extra: int = 4, index: str = "iou"):
    """get_optimal_k(y, Z, 'iou') of hpcs/utils/scores.py:141-177 restated on confusion counts (no sklearn): remap the
    labels to 0..T-1 (:126-139); for k = 1 .. T+extra cut the dendrogram (fcluster maxclust, :151), IoU of every
    (true part, cluster) pair stored in a float32 matrix (:153,158), each true part takes its first best cluster
    (torch.max, :159), later parts overwrite earlier ones in the remap (:161-162), score = agreements / (2N - agreements)
    (one-hot and / or, :164-166); first strictly best k wins.  Returns (pred 0-based or None, k, score)."""
    y = np.asarray(y)
    N = y.shape[0]
    uniq = np.unique(y)
    yt = np.searchsorted(uniq, y)
    T = len(uniq)
    best = (None, 0, 0.0)
    for k in range(1, T + extra + 1):
        yp = fcluster_maxclust_restated(Z, k) - 1
        P = int(yp.max()) + 1
        C = np.zeros((T, P), dtype=np.int64)
        np.add.at(C, (yt, yp), 1)
        ct, cp = C.sum(1), C.sum(0)
        union = ct[:, None] + cp[None, :] - C
        if index == "ri":                                         # scores.py:154-159: adjusted Rand index of the cut
            score = adjusted_rand_from_contingency(C)
            if score > best[2]:
                best = (yp.astype(np.int32), k, score)
            continue
        iou = np.where(union > 0, C / np.maximum(union, 1), 0.0).astype(np.float32)
        ind = iou.argmax(1)                                   # first maximum, like torch.max on CPU
        owner = np.full(P, -1)
        for i in range(T):
            owner[ind[i]] = i                                 # later parts overwrite
        agree = int(sum(C[owner[j], j] for j in range(P) if owner[j] >= 0))
        score = agree / (2 * N - agree)
        if score > best[2]:
            best = (yp.astype(np.int32), k, score)
    return best


# --------------------------------------------------------------------------------------------
# row f-4: the per-step input pipeline (rotation + transpose + one-hot)
#
# PARITY UNPINNED for the rotation matrices: they come from pytorch3d 0.7.2 (hpcs-env.yaml:287), which is neither
# vendored under /root/reference nor installed in this image, so its published algorithm is restated here
# (transforms/rotation_conversions.py: random_quaternions, quaternion_to_matrix; transforms/transform3d.py: Rotate,
# RotateAxisAngle) and anchored on the reference's call sites (hpcs/models/shapenet_hyp_hc.py:63-69,
# partnet_hyp_hc.py:82-88).  Property checks in the tests: orthonormal, det +1, rotation-invariant backbone input norms.
# --------------------------------------------------------------------------------------------
def quaternion_rotations(o: torch.Tensor) -> torch.Tensor:
    """``random_rotations(n)`` given its ``randn(n, 4)`` draws ``o``: unit quaternion with non-negative real part ->
    rotation matrix [n,3,3]."""
    nrm = torch.sqrt((o * o).sum(1))
    nrm = torch.where((nrm < 0) != (o[:, 0] < 0), -nrm, nrm)
    q = o / nrm[:, None]
    r, i, j, k = torch.unbind(q, -1)
    ts = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - ts * (j * j + k * k), ts * (i * j - k * r), ts * (i * k + j * r),
                     ts * (i * j + k * r), 1 - ts * (i * i + k * k), ts * (j * k - i * r),
                     ts * (i * k - j * r), ts * (j * k + i * r), 1 - ts * (i * i + j * j)), -1)
    return m.reshape(-1, 3, 3)


def z_rotations(u: torch.Tensor) -> torch.Tensor:
    """``RotateAxisAngle(angle=u * 360, axis='Z', degrees=True)`` given its ``rand(n)`` draws: the column-vector
    rotation about z, transposed because points are row vectors."""
    a = u * 360 / 180.0 * torch.pi
    c, s, one, zero = torch.cos(a), torch.sin(a), torch.ones_like(a), torch.zeros_like(a)
    R = torch.stack((c, -s, zero, s, c, zero, zero, zero, one), -1).reshape(-1, 3, 3)
    return R.transpose(1, 2)


def rotate_points(points: torch.Tensor, R: Optional[torch.Tensor]) -> torch.Tensor:
    """``trot.transform_points(points)`` then ``.transpose(2, 1)``: points[B,N,3] -> [B,3,N] = (p @ R)^T."""
    out = points if R is None else torch.bmm(points, R.to(points.dtype))
    return out.transpose(2, 1).contiguous()


def to_categorical(y: torch.Tensor, num_classes: int) -> torch.Tensor:
    """hpcs/utils/data.py:24-29."""
    return torch.eye(num_classes)[y.cpu().numpy(),]


# --------------------------------------------------------------------------------------------
# row f-1: the VN layers that consume the edge features (for the fused EdgeConv kernel)
# --------------------------------------------------------------------------------------------
VN_EPS = 1e-6                           # hpcs/nn/dgcnn/utils/vn_layers.py:10


def vn_linear_leaky_relu(x, wf, wd, gamma, beta, running_mean=None, running_var=None, training=True,
                         eps: float = 1e-5, momentum: float = 0.1, negative_slope: float = 0.2):
    """``VNLinearLeakyReLU.forward`` (vn_layers.py:62-77) with its ``VNBatchNorm`` (:118-131, BatchNorm2d on the vector
    norms): x[B,Cin,3,N,k] -> [B,Cout,3,N,k].  ``wf`` / ``wd``: ``map_to_feat`` / ``map_to_dir`` weights [Cout,Cin].  In
    training mode the batch statistics are used and, when given, the running buffers are updated in place like
    ``nn.BatchNorm2d`` (biased variance to normalise, unbiased into ``running_var``)."""
    p = torch.einsum("oi,bicnk->bocnk", wf, x)                                   # :66
    d = torch.einsum("oi,bicnk->bocnk", wd, x)                                   # :70
    norm = torch.sqrt((p * p).sum(2)) + VN_EPS                                   # :124
    if training:
        mean = norm.mean(dim=(0, 2, 3))
        var = norm.var(dim=(0, 2, 3), unbiased=False)
        if running_mean is not None:
            with torch.no_grad():
                m = norm.numel() // norm.shape[1]
                running_mean.mul_(1 - momentum).add_(momentum * mean.detach().to(running_mean.dtype))
                running_var.mul_(1 - momentum).add_(momentum * (var.detach() * m / (m - 1)).to(running_var.dtype))
    else:
        mean, var = running_mean.to(norm.dtype), running_var.to(norm.dtype)
    shape = (1, -1, 1, 1)
    norm_bn = (norm - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * gamma.view(shape) + beta.view(shape)
    p = p / norm.unsqueeze(2) * norm_bn.unsqueeze(2)                             # :128
    dot = (p * d).sum(2, keepdim=True)                                           # :71
    mask = (dot >= 0).to(p.dtype)
    dns = (d * d).sum(2, keepdim=True)
    return negative_slope * p + (1 - negative_slope) * (mask * p + (1 - mask) * (p - (dot / (dns + VN_EPS)) * d))   # :74-76


def edgeconv_layer(x, idx, convs, training=True, k: Optional[int] = None):
    """One graph layer of ``VN_DGCNN_partseg.forward`` (vn_dgcnn_partseg.py:65-68 / 70-73 / 75-77):
    ``get_graph_feature`` -> one or two ``VNLinearLeakyReLU`` -> ``mean_pool`` over k.  ``convs``: list of dicts with keys
    wf, wd, gamma, beta and optionally running_mean / running_var / eps / momentum."""
    e = graph_feature(x, k if k is not None else idx.shape[2], idx=idx)
    for c in convs:
        e = vn_linear_leaky_relu(e, c["wf"], c["wd"], c["gamma"], c["beta"], c.get("running_mean"), c.get("running_var"),
                                 training, c.get("eps", 1e-5), c.get("momentum", 0.1))
    return e.mean(dim=-1)                                                         # vn_layers.py:152-153
