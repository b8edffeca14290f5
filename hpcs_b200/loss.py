"""Triplet mining and the hyperbolic metric-learning objective -- drop-ins for ``hpcs/miner``,
``hpcs/distances/cosine.py`` and ``hpcs/loss/ultrametric_loss.py``.

  get_balanced_random_triplet_indices   hpcs/miner/loss_and_miner_utils.py:7-75
  RandomTripletMarginMiner              hpcs/miner/triplet_margin_miner.py:6-38
  CosineSimilarity                      hpcs/distances/cosine.py:4-16
  MetricHyperbolicLoss                  hpcs/loss/ultrametric_loss.py:16-143

The sampler stays on the host and consumes the torch CPU generator exactly like the reference, so a
common seed gives identical triplets.  Everything downstream of the sampled indices (similarities,
easy/semihard filter, LCA distances, softmax-weighted loss, mean similarity, gradients) is one
fused CUDA pass (csrc/hyp_loss.cu); no (B.N)^2 matrix is built.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import weakref

import torch
import torch.nn.functional as F

from . import _lib

# filter_mode of the C ABI.  The miner-facing names follow hpcs/miner/triplet_margin_miner.py:24-32: 'easy' keeps
# gap > margin; every other value (the constructor default 'all' included) keeps gap <= margin, 'hard' / 'semihard'
# narrow that further.  "none" (mode 0, keep everything) is not a miner type: compute_hyp uses it when miner=False.
_BELOW_MARGIN = 4
FILTER_MODES = {"none": 0, "easy": 1, "semihard": 2, "hard": 3, "all": _BELOW_MARGIN}


# ------------------------------------------------------------------------------------------------
# sampling (host, RNG-parity with the reference)
# ------------------------------------------------------------------------------------------------
def get_balanced_random_triplet_indices(labels: torch.Tensor, ref_labels=None, t_per_anchor: Optional[int] = None,
                                        fraction: Optional[float] = None, weights=None
                                        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per label l (ascending): every member is an anchor ``k_l = int(t_per_anchor * (max_count /
    n_l) ** fraction)`` times, with a positive drawn uniformly from the other members and a negative
    drawn uniformly from the non-members.  Labels with < 2 members or no non-member are skipped."""
    if ref_labels is not None or weights is not None:
        raise NotImplementedError("only the ref_labels=None, weights=None path used by HPCS is provided")
    device = labels.device
    host = labels.detach().cpu()
    counts = torch.bincount(host)
    biggest = counts.max()
    # the draws (torch.randint on the default CPU generator, one pair of calls per label, same sizes, same order) are what makes a
    # patched run consume the RNG stream like the reference; everything around them is laid out to touch each of the T0 triplets
    # once: sizes first, then every label writes its slice of three preallocated (pinned, when bound for the GPU) buffers
    plan = []
    for value in torch.unique(host):
        m = int(counts[value])
        if m < 2 or host.numel() - m < 1:
            continue
        reps = m if t_per_anchor is None else int(t_per_anchor * torch.pow(biggest / m, fraction))
        plan.append((value, m, reps))
    T0 = sum(m * reps for _, m, reps in plan)
    if T0 == 0:
        empty = torch.empty(0, dtype=torch.long, device=device)
        return empty, empty.clone(), empty.clone()
    pin = device.type == "cuda"
    out = [torch.empty(T0, dtype=torch.long, pin_memory=pin) for _ in range(3)]
    at = 0
    for value, m, reps in plan:
        total = m * reps
        inside = host == value
        members = inside.nonzero(as_tuple=True)[0]
        outside = (~inside).nonzero(as_tuple=True)[0]
        pos_draw = torch.randint(0, m - 1, (total,))
        anchor_slot = torch.arange(m).repeat_interleave(reps)
        pos_draw += (pos_draw >= anchor_slot)            # skip the anchor itself
        neg_draw = torch.randint(0, outside.numel(), (total,))
        torch.index_select(members, 0, anchor_slot, out=out[0][at:at + total])
        torch.index_select(members, 0, pos_draw, out=out[1][at:at + total])
        torch.index_select(outside, 0, neg_draw, out=out[2][at:at + total])
        at += total
    return tuple(t.to(device, non_blocking=True) for t in out)


def triplet_segments(label_counts: torch.Tensor, t_per_anchor: Optional[int], fraction: Optional[float]):
    """Host-side plan of the device sampler from the per-label counts (``bincount`` of the labels, a CPU tensor):
    ``seg[4, L]`` int64 = {start in the label-sorted order, members, repeats per anchor, first triplet} for every
    label the reference sampler keeps (>= 2 members, >= 1 non-member), and the total number of triplets."""
    counts = label_counts.detach().cpu().to(torch.int64)
    n = int(counts.sum())
    starts = torch.cumsum(counts, 0) - counts
    ok = (counts >= 2) & (counts < n)
    m = counts[ok]
    if m.numel() == 0:
        return torch.zeros((4, 0), dtype=torch.int64), 0
    if t_per_anchor is None:
        reps = m.clone()
    else:                                                     # int(t * (max_count / m) ** fraction), in torch like the reference
        reps = (t_per_anchor * torch.pow(counts.max() / m, fraction)).to(torch.int64)
    total = m * reps
    first = torch.cumsum(total, 0) - total
    keep = reps > 0
    seg = torch.stack([starts[ok][keep], m[keep], reps[keep], first[keep]]).contiguous()
    return seg, int(total.sum())


def triplet_plan(labels_cpu: torch.Tensor, t_per_anchor: Optional[int], fraction: Optional[float]):
    """Everything the device sampler needs besides a seed, computed from labels that are still on the host (where the
    data loader made them): ``order`` = stable argsort of the labels (int32), ``seg`` and ``T0`` from
    :func:`triplet_segments`.  131 KB + 1 KB to upload for 32768 points, instead of 12 bytes per triplet."""
    lab = labels_cpu.detach().cpu().reshape(-1)
    order = torch.sort(lab, stable=True)[1].to(torch.int32)
    seg, T0 = triplet_segments(torch.bincount(lab), t_per_anchor, fraction)
    return order, seg, T0


def sampler_state(seed: int = 0, device=None) -> torch.Tensor:
    """Device-resident key of the device sampler: int64[3] = {seed, step, 0}.  Every launch that is handed this tensor
    draws from Philox keyed by (seed, step) and then advances ``step`` on the device, so a CUDA graph that captured the
    launch produces fresh triplets on every replay."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0, 0], dtype=torch.int64, device=dev)


def sample_triplets_device(labels: Optional[torch.Tensor], t_per_anchor: Optional[int] = None,
                           fraction: Optional[float] = None, seed: int = 0, plan=None, state: Optional[torch.Tensor] = None
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``get_balanced_random_triplet_indices`` on the GPU -> three int32 index tensors on the device.
    Anchors come out exactly as the reference orders them; positives / negatives are drawn with the same distribution
    from Philox keyed by ``seed`` (not torch's CPU stream).
    ``plan`` = (order, seg, T0) from :func:`triplet_plan`, already on the device: one kernel launch, nothing else
    (capturable in a CUDA graph).  Without a plan it is derived from ``labels`` on the device (a sort, a bincount and
    one device->host copy of the per-label counts)."""
    lib = _lib.load()
    if plan is not None:
        order, seg, T0 = plan
        dev = _lib.require_cuda(order, seg)
    else:
        dev = _lib.require_cuda(labels)
        lab = labels.detach().reshape(-1)
        order = torch.sort(lab, stable=True)[1].to(torch.int32)
        seg, T0 = triplet_segments(torch.bincount(lab).cpu(), t_per_anchor, fraction)       # synchronises
        seg = seg.to(dev)
    out = tuple(torch.empty(T0, dtype=torch.int32, device=dev) for _ in range(3))
    if T0 == 0:
        return out
    if order.dtype != torch.int32 or seg.dtype != torch.int64:
        raise TypeError("plan: order must be int32 and seg int64")
    with torch.cuda.device(dev):
        if state is not None:
            if state.dtype != torch.int64 or state.numel() != 3 or state.device != dev or not state.is_contiguous():
                raise TypeError("state must be the int64[3] device tensor made by sampler_state()")
            _lib.check(lib.hpcs_triplet_sample_state_i32(order.data_ptr(), order.numel(), seg.data_ptr(), seg.shape[1],
                                                         T0, state.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                                         out[2].data_ptr(), _lib.stream_ptr(dev)),
                       "hpcs_triplet_sample_state_i32")
        else:
            _lib.check(lib.hpcs_triplet_sample_i32(order.data_ptr(), order.numel(), seg.data_ptr(), seg.shape[1], T0,
                                                   int(seed) & 0xFFFFFFFFFFFFFFFF, out[0].data_ptr(), out[1].data_ptr(),
                                                   out[2].data_ptr(), _lib.stream_ptr(dev)), "hpcs_triplet_sample_i32")
    return out


# ------------------------------------------------------------------------------------------------
# similarity
# ------------------------------------------------------------------------------------------------
class CosineSimilarity(torch.nn.Module):
    """``0.5 * (1 + q_hat . r_hat)`` on L2-normalised rows; an inverted distance (larger = closer).
    Thin PyTorch: the hot path never calls it (it would build the [n,m] matrix)."""

    is_inverted = True
    normalize_embeddings = True

    def forward(self, query_emb, ref_emb=None):
        q = F.normalize(query_emb, p=2, dim=1)
        r = q if ref_emb is None else F.normalize(ref_emb, p=2, dim=1)
        return self.compute_mat(q, r)

    def compute_mat(self, query_emb, ref_emb):
        return 0.5 * (1 + torch.matmul(query_emb, ref_emb.t()))

    def pairwise_distance(self, query_emb, ref_emb):
        return 0.5 * (1 + torch.sum(query_emb * ref_emb, dim=1))

    def margin(self, x, y):
        return y - x

    def smallest_dist(self, *args, **kwargs):
        return torch.max(*args, **kwargs)


# ------------------------------------------------------------------------------------------------
# fused objective
# ------------------------------------------------------------------------------------------------
def _as_index(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """int64 (what the reference sampler produces) or int32 (half the upload) index tensor on `dev`."""
    dtype = torch.int32 if t.dtype == torch.int32 else torch.int64
    return t.to(device=dev, dtype=dtype).contiguous()


def _same_index_dtype(a, p, n):
    if a.dtype != p.dtype or a.dtype != n.dtype:
        a, p, n = a.long(), p.long(), n.long()
    return a, p, n


def filter_triplets(x: torch.Tensor, a: torch.Tensor, p: torch.Tensor, n: torch.Tensor, margin: float = 0.0,
                    type_of_triplets: str = "easy") -> torch.Tensor:
    """Boolean keep-mask of the miner's margin test (hpcs/miner/triplet_margin_miner.py:16-32)."""
    dev = _lib.require_cuda(x)
    lib = _lib.load()
    xc = x.detach().contiguous().float()
    nrow, D = xc.shape
    a, p, n = _same_index_dtype(_as_index(a, dev), _as_index(p, dev), _as_index(n, dev))
    T0 = a.numel()
    keep = torch.empty(T0, dtype=torch.uint8, device=dev)
    entry = lib.hpcs_triplet_filter_i32_f32 if a.dtype == torch.int32 else lib.hpcs_triplet_filter_f32
    ws = _lib.workspace(lib.hpcs_hyp_triplet_workspace_bytes(nrow, D), dev)
    mode = FILTER_MODES.get(type_of_triplets, _BELOW_MARGIN)
    with torch.cuda.device(dev):
        _lib.check(entry(xc.data_ptr(), nrow, D, a.data_ptr(), p.data_ptr(), n.data_ptr(), T0,
                         mode, float(margin), keep.data_ptr(), ws.data_ptr(), ws.numel(),
                         _lib.stream_ptr(dev)), "hpcs_triplet_filter_f32")
    return keep.bool()


class RandomTripletMarginMiner(torch.nn.Module):
    """Sample class-balanced triplets, keep those passing the margin test on the cosine similarity."""

    def __init__(self, t_per_anchor, fraction, margin=0.2, type_of_triplets="all", distance=None, sampler="reference",
                 **kwargs):
        super().__init__()
        if sampler not in ("reference", "device"):
            raise ValueError("sampler must be 'reference' (host, torch RNG parity) or 'device' (Philox on the GPU)")
        self.t_per_anchor, self.fraction = t_per_anchor, fraction
        self.margin, self.type_of_triplets = margin, type_of_triplets
        self.distance = distance if distance is not None else CosineSimilarity()
        self.sampler, self._state = sampler, None

    def sample(self, labels):
        if self.sampler == "device" and labels.is_cuda:
            if self._state is None or self._state.device != labels.device:     # key (torch's seed, step) lives on the device
                self._state = sampler_state(torch.initial_seed(), labels.device)
            return sample_triplets_device(labels, self.t_per_anchor, self.fraction, state=self._state)
        return get_balanced_random_triplet_indices(labels, t_per_anchor=self.t_per_anchor, fraction=self.fraction)

    def forward(self, embeddings, labels, ref_emb=None, ref_labels=None):
        with torch.no_grad():
            a, p, n = self.sample(labels)
            keep = filter_triplets(embeddings, a, p, n, self.margin, self.type_of_triplets)
            return a[keep], p[keep], n[keep]


class _HypTripletLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, a, p, n, temperature, filter_mode, margin):
        dev = x.device
        lib = _lib.load()
        nrow, D = x.shape
        T0 = a.numel()
        need_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        kept = torch.empty(1, dtype=torch.int64, device=dev)
        ws = _lib.workspace(lib.hpcs_hyp_triplet_workspace_bytes(nrow, D), dev)
        entry = lib.hpcs_hyp_triplet_fwd_i32_f32 if a.dtype == torch.int32 else lib.hpcs_hyp_triplet_fwd_f32
        with torch.cuda.device(dev):
            _lib.check(entry(x.data_ptr(), nrow, D, a.data_ptr(), p.data_ptr(), n.data_ptr(), T0,
                             scale.data_ptr(), float(temperature), int(filter_mode),
                             float(margin), int(need_grad), loss.data_ptr(), kept.data_ptr(),
                             ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                       "hpcs_hyp_triplet_fwd_f32")
        ctx.save_for_backward(x, scale, ws)
        ctx.mark_non_differentiable(kept)
        return loss.reshape(()), kept

    @staticmethod
    def backward(ctx, gloss, _gkept):
        x, scale, ws = ctx.saved_tensors
        dev = x.device
        lib = _lib.load()
        nrow, D = x.shape
        g = gloss.reshape(1).contiguous().float()
        gx = torch.empty_like(x)
        gscale = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.hpcs_hyp_triplet_bwd_f32(g.data_ptr(), x.data_ptr(), nrow, D, scale.data_ptr(), ws.data_ptr(),
                                                    ws.numel(), gx.data_ptr(), gscale.data_ptr(), _lib.stream_ptr(dev)),
                       "hpcs_hyp_triplet_bwd_f32")
        return gx, gscale.reshape(scale.shape), None, None, None, None, None, None


def hyp_triplet_loss(x: torch.Tensor, triplets, scale: torch.Tensor, temperature: float,
                     type_of_triplets: str = "none", margin: float = 0.0, return_kept: bool = False):
    """``mean_T(total) + mean(mat_sim)`` of ``compute_hyp`` for given (sampled, unfiltered) triplets.
    ``type_of_triplets`` applies the miner's filter inside the kernel ("none": every triplet counts)."""
    dev = _lib.require_cuda(x, scale)
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError("hyp_triplet_loss expects x[n,D] float32")
    a, p, n = _same_index_dtype(*(_as_index(t, dev) for t in triplets))
    sc = scale.reshape(-1)
    if sc.numel() != 1 or sc.dtype != torch.float32:
        raise TypeError("scale must hold one float32")
    mode = FILTER_MODES.get(type_of_triplets, _BELOW_MARGIN)
    loss, kept = _HypTripletLoss.apply(x.contiguous(), sc if sc.is_contiguous() else sc.contiguous(), a, p, n,
                                       temperature, mode, margin)
    return (loss, kept) if return_kept else loss


class CosFaceLoss(torch.nn.Module):
    """Large-margin cosine loss, ``logits = s (cos - m onehot)`` + mean cross entropy: the arithmetic
    the reference takes from pytorch-metric-learning (visible at
    hpcs/loss/hierarchical_cosface_loss.py:57-74).  Outside the hot path; plain PyTorch."""

    def __init__(self, num_classes, embedding_size, margin=0.35, scale=64):
        super().__init__()
        self.margin, self.scale = margin, scale
        self.W = torch.nn.Parameter(torch.empty(embedding_size, num_classes))
        torch.nn.init.normal_(self.W)

    def get_cosine(self, embeddings):
        return F.normalize(embeddings, p=2, dim=1) @ F.normalize(self.W, p=2, dim=0)

    def get_logits(self, embeddings, labels):
        cosine = self.get_cosine(embeddings)
        onehot = F.one_hot(labels.long(), cosine.shape[1]).to(cosine.dtype)
        return (cosine - self.margin * onehot) * self.scale

    def forward(self, embeddings, labels):
        return F.cross_entropy(self.get_logits(embeddings, labels), labels.long())


def triplet_margin_loss(x: torch.Tensor, a, p, n, margin: float, distance=None) -> torch.Tensor:
    """The ``--triplet-sim`` metric term: ``relu(sim(a,n) - sim(a,p) + margin)`` averaged over the violating triplets
    (hpcs/miner/triplet_margin_loss.py:34-65 with the inverted cosine similarity and PML's AvgNonZeroReducer), from row
    gathers instead of the [n,n] matrix.  0 (with a graph) when nothing is mined or nothing violates."""
    distance = distance if distance is not None else CosineSimilarity()
    if a.numel() == 0:
        return x.sum() * 0
    u = F.normalize(x, p=2, dim=1)
    ap = distance.pairwise_distance(u[a], u[p])
    an = distance.pairwise_distance(u[a], u[n])
    viol = F.relu(distance.margin(ap, an) + margin)
    positive = viol[viol > 0]
    return positive.mean() if positive.numel() else x.sum() * 0


class MetricHyperbolicLoss(torch.nn.Module):
    """Same constructor and methods as the reference class (ultrametric_loss.py:16-143)."""

    def __init__(self, margin: float = 1.0, t_per_anchor: int = 50, fraction: float = 1.2,
                 scale: Union[float, torch.Tensor, torch.nn.Parameter] = 1e-3, temperature: float = 0.05,
                 anneal_factor: float = 0.5, num_class: int = 4, embedding_size: int = 4, cosface: bool = True,
                 miner: bool = False, sampler: str = "reference"):
        super().__init__()
        self.margin, self.t_per_anchor, self.fraction = margin, t_per_anchor, fraction
        self.scale = scale if isinstance(scale, torch.Tensor) else torch.tensor([float(scale)])
        self.temperature, self.anneal_factor = temperature, anneal_factor
        self.num_class, self.embedding_size = num_class, embedding_size
        self.cosface, self.miner = cosface, miner
        self.distance_sim = CosineSimilarity()
        if self.miner:
            self.hyp_miner = RandomTripletMarginMiner(distance=self.distance_sim, margin=0, t_per_anchor=t_per_anchor,
                                                      fraction=fraction, type_of_triplets="easy", sampler=sampler)
        if self.cosface:
            self.loss_cosface = CosFaceLoss(num_classes=num_class, embedding_size=embedding_size, margin=0.35, scale=2)
        else:
            self.triplet_miner = RandomTripletMarginMiner(distance=self.distance_sim, margin=margin,
                                                          t_per_anchor=t_per_anchor, fraction=fraction,
                                                          type_of_triplets="semihard")

    # -- helpers kept from the reference -----------------------------------------------------------
    def get_triplets(self, n_samples):
        """All pairs (i<j), each ``t_per_anchor`` times, with a random third point != i, j
        (ultrametric_loss.py:42-55; only practical for small n)."""
        pairs = torch.combinations(torch.arange(n_samples), r=2).repeat_interleave(self.t_per_anchor, dim=0)
        third = torch.randint(n_samples, (pairs.shape[0],), dtype=torch.long)
        ok = (pairs[:, 0] != third) & (pairs[:, 1] != third)
        return pairs[ok, 0], pairs[ok, 1], third[ok]

    def _scale_on(self, dev):
        if self.scale.device != dev:
            if isinstance(self.scale, torch.nn.Parameter):
                raise RuntimeError("scale Parameter lives on another device than the embeddings")
            self.scale = self.scale.to(dev)
        return self.scale

    def normalize_embeddings(self, embeddings):
        return F.normalize(embeddings, p=2, dim=1) * torch.clamp(self._scale_on(embeddings.device), 1e-4, 1)

    def compute_hyp(self, x_poincare, labels, triplets=None):
        """HypHC triplet loss + mean similarity.  ``triplets`` (a, p, n) may be passed to reuse an
        already sampled set; otherwise they are sampled like the reference does."""
        dev = x_poincare.device
        if triplets is None:
            triplets = self.hyp_miner.sample(labels) if self.miner else self.get_triplets(x_poincare.shape[0])
        kind, margin = (self.hyp_miner.type_of_triplets, self.hyp_miner.margin) if self.miner else ("none", 0.0)
        return hyp_triplet_loss(x_poincare, triplets, self._scale_on(dev), self.temperature, kind, margin)

    def get_logits(self, embeddings, labels):
        if not hasattr(self, "loss_cosface"):
            raise ValueError("Cannot get logits since this class doesn't use any CosFaceLoss")
        return self.loss_cosface.get_logits(embeddings, labels)

    def compute_loss(self, x_euclidean, x_poincare, labels, *args):
        loss_hyperbolic = self.compute_hyp(x_poincare, labels)
        if self.cosface:
            loss_metric = self.loss_cosface(x_poincare, labels.long())
        else:
            a, p, n = self.triplet_miner(x_poincare, labels)
            loss_metric = triplet_margin_loss(x_poincare, a.long(), p.long(), n.long(), self.margin, self.distance_sim)
        return {"loss_hyp": {"losses": loss_hyperbolic}, "loss_metric": {"losses": loss_metric}}

    def anneal_temperature(self):
        self.temperature *= min(max(self.anneal_factor, 0.2), 1.0)
        return self.temperature


# ------------------------------------------------------------------------------------------------
# PartNet default: hierarchical CosFace on top of the same hyperbolic objective
# ------------------------------------------------------------------------------------------------
def hierarchical_loss(probabilities: torch.Tensor, targets: torch.Tensor, hierarchy_list) -> torch.Tensor:
    """Sum over hierarchy levels of ``nll(log p_level, targets)`` where, level by level, the probability of every
    channel of a branch is replaced by the branch's total (hpcs/loss/hierarchical_cosface_loss.py:9-28).  Branches
    are applied one after another on the same tensor like the reference does, so an (unusual) overlap between two
    branches of one level sums already-summed channels exactly as there."""
    loss = probabilities.new_zeros(())
    for level in hierarchy_list:
        summed = probabilities
        for branch in level:
            cols = torch.as_tensor(list(branch), dtype=torch.long, device=probabilities.device)
            if cols.numel() == 0:
                continue
            total = summed.index_select(1, cols).sum(1, keepdim=True)
            summed = summed.index_copy(1, cols, total.expand(-1, cols.numel()))
        loss = loss + F.nll_loss(torch.log(summed), targets.long())
    return loss


class HierarchicalCosFaceLoss(CosFaceLoss):
    """CosFace logits -> softmax -> :func:`hierarchical_loss` (hierarchical_cosface_loss.py:31-87).  Outside the hot
    path; plain PyTorch, kept so the PartNet default configuration constructs and trains."""

    def __init__(self, num_classes, embedding_size, margin=0.35, scale=64, hierarchy_list=None):
        super().__init__(num_classes, embedding_size, margin=margin, scale=scale)
        self.hierarchy_list = hierarchy_list if hierarchy_list is not None else []

    def forward(self, embeddings, labels):
        probabilities = F.softmax(self.get_logits(embeddings, labels), dim=1)
        return hierarchical_loss(probabilities, labels, self.hierarchy_list)


class HierarchicalMetricHyperbolicLoss(MetricHyperbolicLoss):
    """Same constructor and methods as the reference class (ultrametric_loss.py:146-176): the hyperbolic term is the
    fused CUDA objective of the parent, the metric term the hierarchical CosFace."""

    def __init__(self, margin: float = 1.0, t_per_anchor: int = 50, fraction: float = 1.2,
                 scale: Union[float, torch.Tensor, torch.nn.Parameter] = 1e-3, temperature: float = 0.05,
                 anneal_factor: float = 0.5, num_class: int = 4, embedding_size: int = 4, miner: bool = False,
                 hierarchy_list: Optional[list] = None, sampler: str = "reference"):
        super().__init__(margin=margin, t_per_anchor=t_per_anchor, fraction=fraction, scale=scale,
                         temperature=temperature, anneal_factor=anneal_factor, num_class=num_class,
                         embedding_size=embedding_size, cosface=True, miner=miner, sampler=sampler)
        self.hierarchy_list = hierarchy_list if hierarchy_list is not None else []
        self.loss_cosface = HierarchicalCosFaceLoss(num_classes=num_class, embedding_size=embedding_size, margin=0.35,
                                                    scale=2, hierarchy_list=self.hierarchy_list)

    def compute_loss(self, x_euclidean, x_poincare, labels, *args):
        loss_hyperbolic = self.compute_hyp(x_poincare, labels)
        loss_metric = self.loss_cosface(x_poincare, labels.long())
        return {"loss_hyp": {"losses": loss_hyperbolic}, "loss_metric": {"losses": loss_metric}}


# ------------------------------------------------------------------------------------------------
# methods bound onto the REFERENCE's own classes by hpcs_b200.patch (attribute names are the reference's)
# ------------------------------------------------------------------------------------------------
def cosface_logits(embeddings: torch.Tensor, W: torch.Tensor, labels: torch.Tensor, margin: float, scale: float) -> torch.Tensor:
    """``scale * (cos(emb_i, W_c) - margin * [labels_i == c])`` in one kernel, forward only (no autograd graph): the logits the
    reference's metrics are computed from (hpcs/loss/ultrametric_loss.py:95-112)."""
    dev = _lib.require_cuda(embeddings, W, labels)
    lib = _lib.load()
    e = embeddings.detach().contiguous().float()
    w = W.detach().contiguous().float()
    if e.dim() != 2 or w.dim() != 2 or w.shape[0] != e.shape[1]:
        raise ValueError(f"cosface_logits: embeddings {tuple(e.shape)} vs W {tuple(w.shape)}")
    y = labels.detach().reshape(-1).contiguous().long()
    if y.numel() != e.shape[0]:
        raise ValueError("cosface_logits: one label per embedding row")
    out = torch.empty(e.shape[0], w.shape[1], device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.hpcs_cosface_logits_f32(e.data_ptr(), w.data_ptr(), y.data_ptr(), e.shape[0], e.shape[1], w.shape[1],
                                               float(margin), float(scale), out.data_ptr(), _lib.stream_ptr(dev)),
                   "hpcs_cosface_logits_f32")
    return out


def native_get_logits(self, embeddings, labels):
    """Bound onto ``MetricHyperbolicLoss.get_logits`` (hpcs/loss/ultrametric_loss.py:95-112).  The reference calls it twice per
    step with the same tensors (accuracy and IoU, hpcs/models/base_hyp_hc.py:88-99) and each call runs a matmul, a one-hot mask
    and a boolean-mask gather that synchronises the host; here: one kernel, no synchronisation, and the second call of a step
    returns the first one's result (same tensor objects, same versions, same weight).  Forward only: the callers are metrics."""
    if not hasattr(self, "loss_cosface"):
        raise ValueError("Cannot get logits since this class doesn't use any CosFaceLoss")
    lc = self.loss_cosface
    if hasattr(lc, "cast_types"):
        lc.cast_types(embeddings.dtype, embeddings.device)
    W = lc.W
    key = (embeddings._version, labels._version, W._version, W.data_ptr())
    cached = getattr(self, "_hpcs_logits_cache", None)
    if cached is not None and cached[0]() is embeddings and cached[1]() is labels and cached[2] == key:
        return cached[3]
    logits = cosface_logits(embeddings, W, labels, lc.margin, lc.scale)
    object.__setattr__(self, "_hpcs_logits_cache", (weakref.ref(embeddings), weakref.ref(labels), key, logits))
    return logits


# Which sampler the BOUND methods use (hpcs_b200.patch.install(sampler=...)): "reference" = the reference's sampler on the host,
# torch RNG stream and draws identical to an unpatched run (one device->host copy of the labels, three index uploads per step);
# "device" = the Philox sampler on the GPU (row f-3): same anchors in the same order, positives / negatives from the same
# distributions but a different stream, no host work, key derived from torch.initial_seed().
BOUND_SAMPLER = "reference"


def _bound_sample(owner, labels, t_per_anchor, fraction):
    if BOUND_SAMPLER == "device" and labels.is_cuda:
        st = getattr(owner, "_hpcs_sampler_state", None)
        if st is None or st.device != labels.device:
            st = sampler_state(torch.initial_seed(), labels.device)
            object.__setattr__(owner, "_hpcs_sampler_state", st)
        return sample_triplets_device(labels, t_per_anchor, fraction, state=st)
    return get_balanced_random_triplet_indices(labels, t_per_anchor=t_per_anchor, fraction=fraction)


def native_mine(self, embeddings, labels, ref_emb=None, ref_labels=None):
    """``RandomTripletMarginMiner.mine`` (hpcs/miner/triplet_margin_miner.py:13-38) without the [n,n] matrix: the
    reference's sampler order (and, with the default sampler, its RNG draws), the margin test by ``hpcs_triplet_filter_f32``."""
    a, p, n = _bound_sample(self, labels, self.t_per_anchor, self.fraction)
    keep = filter_triplets(embeddings, a, p, n, self.margin, self.type_of_triplets)
    return a[keep].long(), p[keep].long(), n[keep].long()


def native_compute_hyp(self, x_poincare, labels):
    """``MetricHyperbolicLoss.compute_hyp`` (hpcs/loss/ultrametric_loss.py:57-93) as one fused CUDA pass; also what
    ``HierarchicalMetricHyperbolicLoss`` inherits.  Reads the reference object's own attributes (``miner``,
    ``hyp_miner``, ``scale``, ``temperature``) and never calls ``self.hyp_miner`` / ``self.distance_sim``."""
    if self.miner:
        m = self.hyp_miner
        triplets = _bound_sample(self, labels, m.t_per_anchor, m.fraction)
        kind, margin = m.type_of_triplets, m.margin
    else:
        triplets, kind, margin = self.get_triplets(x_poincare.shape[0]), "none", 0.0
    scale = self.scale if isinstance(self.scale, torch.Tensor) else torch.tensor([float(self.scale)])
    return hyp_triplet_loss(x_poincare, triplets, scale.to(x_poincare.device), float(self.temperature), kind, margin)
