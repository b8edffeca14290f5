"""Install the native path into an importable copy of the reference, so its ``train.py`` /
``infer.py`` run unmodified.  Usage, at any point before the first batch::

    import hpcs_b200.patch; hpcs_b200.patch.install()

Nothing of the reference is edited on disk.  Two kinds of rebinding, chosen so that neither import order nor
subclassing can leave a call site on the reference's PyTorch path:

* **functions** (``knn``, ``get_graph_feature``, ``hyp_lca``, ``get_optimal_k`` ...): the name is rebound in the module
  that defines it AND in every loaded module that holds the same object under any name (``from x import f`` aliases in
  ``vn_dgcnn_partseg``, ``vn_pointnet_partseg``, ``vn_dgcnn_expo``, ``base_hyp_hc``, ``__main__`` ...).
* **methods** of the reference's own classes (``VN_DGCNN_partseg.forward``: the three graph layers fused, row f-1;
  ``MetricHyperbolicLoss.compute_hyp``,
  ``RandomTripletMarginMiner.mine``, ``ExpMap.forward``, ``MLPExpMap.forward``,
  ``BaseSimilarityHypHC._decode_linkage`` and ``.forward`` (batched decode instead of the per-cloud loop),
  ``ShapeNetHypHC._forward``, ``PartNetHypHC._forward``): the class object
  stays the reference's, so subclasses (``HierarchicalMetricHyperbolicLoss``, the PartNet default), ``isinstance``
  checks, checkpoints and hyper-parameter pickles keep working, and instances created BEFORE ``install()`` switch too.

``install()`` raises if any target cannot be bound (``strict=False`` downgrades that to a warning listing what is still
on the reference path); ``verify()`` re-checks the binding at any later time.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings
from typing import Dict, List, Tuple

from . import decode, edgeconv, graph, hyperbolic, loss, pipeline

# (defining module, function name) -> replacement
FUNCTIONS: Dict[Tuple[str, str], object] = {
    ("hpcs.nn.dgcnn.utils.vn_dgcnn_util", "knn"): graph.knn,
    ("hpcs.nn.dgcnn.utils.vn_dgcnn_util", "get_graph_feature"): graph.get_graph_feature,
    ("hpcs.nn.dgcnn.utils.vn_dgcnn_util", "get_graph_feature_cross"): graph.get_graph_feature_cross,
    ("hpcs.nn.pointnet.utils.vn_dgcnn_util", "knn"): graph.knn,
    ("hpcs.nn.pointnet.utils.vn_dgcnn_util", "get_graph_feature"): graph.get_graph_feature,
    ("hpcs.nn.pointnet.utils.vn_dgcnn_util", "get_graph_feature_cross"): graph.get_graph_feature_cross,
    ("hpcs.nn.dgcnn.utils.dgcnn_util", "knn"): graph.knn,                      # same body, [B,C,N] input
    ("hpcs.distances.lca", "hyp_lca"): hyperbolic.hyp_lca,
    ("hpcs.miner.loss_and_miner_utils", "get_balanced_random_triplet_indices"): loss.get_balanced_random_triplet_indices,
    ("hpcs.utils.scores", "get_optimal_k"): decode.get_optimal_k,
}


def _decode_linkage(self, leaves_embeddings):
    """Bound onto ``BaseSimilarityHypHC`` (hpcs/models/base_hyp_hc.py:81-86)."""
    return decode.decode_linkage(leaves_embeddings, self.metric_hyp_loss.scale, method="complete")


def _model_forward(self, batch, testing: bool = False):
    """Bound onto ``BaseSimilarityHypHC.forward`` (hpcs/models/base_hyp_hc.py:120-140): same calls, same return tuples; the
    per-cloud Python loop of ``_decode_linkage`` + device-to-host copy (:133-137) is one batched decode, with Z copied to
    the host once (the list of per-cloud numpy matrices the caller expects)."""
    points, x_euclidean, x_poincare, pts_labels = self._forward(batch, testing)
    xe = x_euclidean.contiguous().view(-1, x_euclidean.shape[-1])
    xp = x_poincare.contiguous().view(-1, x_poincare.shape[-1])
    losses = self.compute_losses(xe, xp, pts_labels)
    metrics = {}
    if hasattr(self.metric_hyp_loss, "loss_cosface"):
        y_true = pts_labels.contiguous().reshape(-1).long()
        metrics = {"acc": self.compute_accuracy(xp, y_true), "iou": self.compute_iou(xp, y_true)}
    if not testing:
        return losses, metrics
    Z = decode.decode_linkage_batch(x_poincare, self.metric_hyp_loss.scale, method="complete")
    return losses, metrics, x_euclidean, x_poincare, list(Z.cpu().numpy()), points, pts_labels


def _expmap_forward(self, x):
    """Bound onto ``ExpMap`` (hpcs/nn/hyperbolic/hyp_embed.py:9-10): ``expmap_1(x, 0)``."""
    return hyperbolic.expmap0(x)


def _mlp_expmap_forward(self, x):
    """Bound onto ``MLPExpMap`` (hyp_embed.py:21-23): the Linear stays PyTorch, the map is the CUDA op."""
    return hyperbolic.expmap0(self.mlp(x))


def _triplet_margin_compute_loss(self, embeddings, labels, indices_tuple, ref_emb=None, ref_labels=None):
    """Bound onto the reference's ``TripletMarginLoss`` (hpcs/miner/triplet_margin_loss.py:34-62, the ``--triplet-sim``
    variant): same loss dictionary, but the three similarities per mined triplet come from row gathers instead of the
    [n,n] matrix ``self.distance(embeddings, ref_emb)``.  Anything but an explicit (a, p, n) tuple goes to the original."""
    import torch.nn.functional as F
    if indices_tuple is None or len(indices_tuple) != 3 or (ref_emb is not None and ref_emb is not embeddings):
        return _ORIGINALS[("hpcs.miner.triplet_margin_loss", "TripletMarginLoss", "compute_loss")](
            self, embeddings, labels, indices_tuple, ref_emb, ref_labels)
    a, p, n = indices_tuple
    if len(a) == 0:
        return self.zero_losses()
    u = F.normalize(embeddings, p=2, dim=1)
    ap = self.distance.pairwise_distance(u[a], u[p])
    an = self.distance.pairwise_distance(u[a], u[n])
    if self.swap:
        an = self.distance.smallest_dist(an, self.distance.pairwise_distance(u[p], u[n]))
    violation = self.distance.margin(ap, an) + self.margin
    losses = F.softplus(violation) if self.smooth_loss else F.relu(violation)
    return {"loss": {"losses": losses, "indices": indices_tuple, "reduction_type": "triplet"}}


# (defining module, class name, method name) -> replacement
METHODS: Dict[Tuple[str, str, str], object] = {
    ("hpcs.loss.ultrametric_loss", "MetricHyperbolicLoss", "compute_hyp"): loss.native_compute_hyp,
    ("hpcs.loss.ultrametric_loss", "MetricHyperbolicLoss", "get_logits"): loss.native_get_logits,
    ("hpcs.miner.triplet_margin_miner", "RandomTripletMarginMiner", "mine"): loss.native_mine,
    ("hpcs.miner.triplet_margin_loss", "TripletMarginLoss", "compute_loss"): _triplet_margin_compute_loss,
    ("hpcs.nn.hyperbolic.hyp_embed", "ExpMap", "forward"): _expmap_forward,
    ("hpcs.nn.hyperbolic.hyp_embed", "MLPExpMap", "forward"): _mlp_expmap_forward,
    ("hpcs.nn.dgcnn.vn_dgcnn_partseg", "VN_DGCNN_partseg", "forward"): edgeconv.vn_dgcnn_partseg_forward,
    ("hpcs.models.base_hyp_hc", "BaseSimilarityHypHC", "_decode_linkage"): _decode_linkage,
    ("hpcs.models.base_hyp_hc", "BaseSimilarityHypHC", "forward"): _model_forward,
    ("hpcs.models.shapenet_hyp_hc", "ShapeNetHypHC", "_forward"): pipeline.shapenet_forward,
    ("hpcs.models.partnet_hyp_hc", "PartNetHypHC", "_forward"): pipeline.partnet_forward,
}

_ORIGINALS: Dict[Tuple[str, ...], object] = {}      # what each target held before install(), for uninstall()/verify()
_ALIASES: Dict[Tuple[str, str], List[str]] = {}     # per replaced function: the alias names rebound besides the definition


class PatchError(RuntimeError):
    pass


def _module(name: str):
    return sys.modules.get(name) or importlib.import_module(name)


def _rebind_aliases(original, replacement) -> List[str]:
    """Every attribute of every loaded module that IS ``original`` -> ``replacement``."""
    done = []
    for mod_name, mod in list(sys.modules.items()):
        if mod is None or mod_name.startswith("hpcs_b200"):
            continue
        try:
            names = [k for k, v in vars(mod).items() if v is original]
        except Exception:
            continue
        for k in names:
            setattr(mod, k, replacement)
            done.append(f"{mod_name}.{k}")
    return done


def install(strict: bool = True, sampler: str = None) -> List[str]:
    """Bind every target; returns the fully qualified names now on the native path.  A target that cannot be bound
    raises :class:`PatchError` (``strict``) or is reported by one warning.
    ``sampler``: which triplet sampler the bound loss / miner methods use -- "reference" (default: the reference's host sampler,
    identical torch RNG draws) or "device" (Philox on the GPU, no host work per step, distributionally equal); None reads the
    environment variable HPCS_B200_SAMPLER, else "reference"."""
    if sampler is None:
        sampler = os.environ.get("HPCS_B200_SAMPLER", "reference")
    if sampler not in ("reference", "device"):
        raise ValueError("sampler must be 'reference' or 'device'")
    loss.BOUND_SAMPLER = sampler
    done, failed = [], []
    for (mod_name, attr), repl in FUNCTIONS.items():
        try:
            mod = _module(mod_name)
            cur = getattr(mod, attr)
        except Exception as e:                               # noqa: BLE001 - reported below
            failed.append(f"{mod_name}.{attr}: {type(e).__name__}: {e}")
            continue
        if cur is not repl:
            _ORIGINALS.setdefault((mod_name, attr), cur)
            setattr(mod, attr, repl)
        orig = _ORIGINALS.get((mod_name, attr))
        done.append(f"{mod_name}.{attr}")
        if orig is not None:
            rebound = _rebind_aliases(orig, repl)
            _ALIASES.setdefault((mod_name, attr), []).extend(rebound)
            done += [n for n in rebound if n not in done]
    for (mod_name, cls_name, meth), repl in METHODS.items():
        try:
            cls = getattr(_module(mod_name), cls_name)
            cur = cls.__dict__.get(meth)
            if cur is None and not hasattr(cls, meth):
                raise AttributeError(f"{cls_name} has no method {meth}")
        except Exception as e:                               # noqa: BLE001
            failed.append(f"{mod_name}.{cls_name}.{meth}: {type(e).__name__}: {e}")
            continue
        if cur is not repl:
            _ORIGINALS.setdefault((mod_name, cls_name, meth), cur)
            setattr(cls, meth, repl)
            if meth == "forward" and cls_name == "VN_DGCNN_partseg":          # pooling='max' keeps the module's own layer sequence
                cls._hpcs_reference_forward = _ORIGINALS[(mod_name, cls_name, meth)]
        done.append(f"{mod_name}.{cls_name}.{meth}")
    if failed:
        msg = ("hpcs_b200.patch.install(): these call sites are still on the reference's PyTorch path:\n  "
               + "\n  ".join(failed))
        if strict:
            raise PatchError(msg)
        warnings.warn(msg, RuntimeWarning, stacklevel=2)
    return done


def verify() -> List[str]:
    """Names that are NOT on the native path right now (empty list = fully bound).  Checks the defining modules, every
    alias of a replaced function in any loaded module, and the method tables of the reference's classes, including
    what subclasses resolve through their MRO."""
    missing = []
    for (mod_name, attr), repl in FUNCTIONS.items():
        mod = sys.modules.get(mod_name)
        if mod is None or getattr(mod, attr, None) is not repl:
            missing.append(f"{mod_name}.{attr}")
        orig = _ORIGINALS.get((mod_name, attr))
        if orig is None:
            continue
        for other_name, other in list(sys.modules.items()):
            if other is None or other_name.startswith("hpcs_b200"):
                continue
            try:
                missing += [f"{other_name}.{k}" for k, v in vars(other).items() if v is orig]
            except Exception:
                continue
    for (mod_name, cls_name, meth), repl in METHODS.items():
        mod = sys.modules.get(mod_name)
        cls = getattr(mod, cls_name, None) if mod is not None else None
        if cls is None or cls.__dict__.get(meth) is not repl:
            missing.append(f"{mod_name}.{cls_name}.{meth}")
            continue
        stack = list(cls.__subclasses__())
        while stack:                                         # a subclass that overrides the method is reported
            sub = stack.pop()
            if getattr(sub, meth) is not repl:
                missing.append(f"{sub.__module__}.{sub.__qualname__}.{meth} (overrides the patched method)")
            stack += sub.__subclasses__()
    return missing


def uninstall() -> None:
    """Put back everything :func:`install` replaced (A/B runs against the reference path)."""
    for key, orig in list(_ORIGINALS.items()):
        if len(key) == 2:
            mod_name, attr = key
            for alias in _ALIASES.pop(key, []):
                owner, _, name = alias.rpartition(".")
                if owner in sys.modules:
                    setattr(sys.modules[owner], name, orig)
            setattr(sys.modules[mod_name], attr, orig)
        else:
            mod_name, cls_name, meth = key
            cls = getattr(sys.modules[mod_name], cls_name)
            if orig is None:
                delattr(cls, meth)
            else:
                setattr(cls, meth, orig)
        del _ORIGINALS[key]
