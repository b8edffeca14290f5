"""Import shims that let the UNMODIFIED reference (``/root/reference``) run in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (and nothing else) to execute the
reference's own code on CPU and record golden vectors under ``tests/golden/``.

The reference depends on ``pytorch-metric-learning`` 1.6.3 (``hpcs-env.yaml:283``), which is not
installed here and cannot be (no network).  Only a handful of its base classes are touched on the
hot path; their published behaviour is restated below (nothing is copied from the reference):

* ``BaseDistance.__call__``: L2-normalise rows (``F.normalize``, p=2, eps 1e-12) when
  ``normalize_embeddings`` is set, call ``compute_mat``, raise to ``power`` if != 1.
  Call sites: ``hpcs/distances/cosine.py:4-16``, ``hpcs/miner/triplet_margin_miner.py:16``,
  ``hpcs/loss/ultrametric_loss.py:65``.
* ``DotProductSimilarity``: ``is_inverted=True``; ``compute_mat = q @ r.T``.
* ``BaseMiner.forward``: run ``mine`` under ``no_grad`` with ``ref_emb = embeddings`` when no
  reference set is given.  Call site: ``hpcs/miner/triplet_margin_miner.py:6-13``.
* ``BaseMetricLossFunction`` / ``CosFaceLoss`` / ``LargeMarginSoftmaxLoss`` / reducers: only need
  to be constructible so that ``hpcs.loss.ultrametric_loss`` imports; the CosFace term is outside
  the hot path and is not used for goldens.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn.functional as F

REFERENCE_ROOT = "/root/reference"


def _module(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


class _BaseDistance(torch.nn.Module):
    def __init__(self, normalize_embeddings=True, p=2, power=1, is_inverted=False, **kwargs):
        super().__init__()
        self.normalize_embeddings = normalize_embeddings
        self.p = p
        self.power = power
        self.is_inverted = is_inverted

    def forward(self, query_emb, ref_emb=None):
        q = self.maybe_normalize(query_emb)
        r = q if ref_emb is None else self.maybe_normalize(ref_emb)
        mat = self.compute_mat(q, r)
        if self.power != 1:
            mat = mat ** self.power
        return mat

    def normalize(self, embeddings, dim=1, **kwargs):
        return F.normalize(embeddings, p=self.p, dim=dim, **kwargs)

    def maybe_normalize(self, embeddings, dim=1, **kwargs):
        if self.normalize_embeddings:
            return self.normalize(embeddings, dim=dim, **kwargs)
        return embeddings

    def smallest_dist(self, *args, **kwargs):
        return torch.max(*args, **kwargs) if self.is_inverted else torch.min(*args, **kwargs)

    def margin(self, x, y):
        return y - x if self.is_inverted else x - y


class _DotProductSimilarity(_BaseDistance):
    def __init__(self, **kwargs):
        super().__init__(is_inverted=True, **kwargs)

    def compute_mat(self, query_emb, ref_emb):
        return torch.matmul(query_emb, ref_emb.t())

    def pairwise_distance(self, query_emb, ref_emb):
        return torch.sum(query_emb * ref_emb, dim=1)


class _BaseMiner(torch.nn.Module):
    def __init__(self, distance=None, **kwargs):
        super().__init__()
        self.distance = distance

    def forward(self, embeddings, labels, ref_emb=None, ref_labels=None):
        with torch.no_grad():
            labels = labels.to(embeddings.device)
            if ref_emb is None:
                ref_emb, ref_labels = embeddings, labels
            return self.mine(embeddings, labels, ref_emb, ref_labels)


class _TripletMarginMiner(_BaseMiner):
    def __init__(self, margin=0.2, type_of_triplets="all", **kwargs):
        super().__init__(**kwargs)
        self.margin = margin
        self.type_of_triplets = type_of_triplets


class _BaseMetricLossFunction(torch.nn.Module):
    """PML 1.6.3 ``BaseMetricLossFunction`` as far as the reference's losses use it: ``forward`` calls
    ``compute_loss(embeddings, labels, indices_tuple, ref_emb=embeddings, ref_labels=labels)`` and reduces the
    ``"loss"`` entry -- mean over elements (MeanReducer), or mean over the positive entries for a class whose
    ``get_default_reducer`` names ``AvgNonZeroReducer`` (0 when there is none)."""

    def __init__(self, distance=None, **kwargs):
        super().__init__()
        self.distance = distance

    def add_to_recordable_attributes(self, *args, **kwargs):
        pass

    def zero_losses(self):
        return {"loss": {"losses": 0, "indices": None, "reduction_type": "already_reduced"}}

    def forward(self, embeddings, labels=None, indices_tuple=None):
        entry = self.compute_loss(embeddings, labels, indices_tuple, embeddings, labels)["loss"]
        losses = entry["losses"]
        if entry["reduction_type"] == "already_reduced" or not torch.is_tensor(losses):
            return embeddings.sum() * 0 + losses
        if hasattr(self, "get_default_reducer") and isinstance(self.get_default_reducer(), _Reducer):
            nz = losses[losses > 0]
            return nz.mean() if nz.numel() else embeddings.sum() * 0
        return losses.mean()


class _LargeMarginSoftmaxLoss(_BaseMetricLossFunction):
    """PML 1.6.3 ``LargeMarginSoftmaxLoss`` members that ``hpcs/loss/hierarchical_cosface_loss.py`` and
    ``MetricHyperbolicLoss.get_logits`` call: ``W[embedding_size, num_classes]`` normal-initialised, cosine of the
    L2-normalised embeddings and columns, one-hot target mask."""
    collect_stats = False

    def __init__(self, num_classes=2, embedding_size=2, margin=4, scale=1, **kwargs):
        super().__init__(**kwargs)
        self.margin, self.scale, self.num_classes = margin, scale, num_classes
        self.W = torch.nn.Parameter(torch.randn(embedding_size, num_classes))

    def get_cosine(self, embeddings):
        return F.normalize(embeddings, p=2, dim=1) @ F.normalize(self.W, p=2, dim=0)

    def get_target_mask(self, embeddings, labels):
        return F.one_hot(labels.long(), self.W.shape[1]).to(embeddings.dtype)

    def cast_types(self, dtype, device):
        self.W.data = self.W.data.to(device=device, dtype=dtype)

    def add_weight_regularization_to_loss_dict(self, loss_dict, weights):
        pass


class _CosFaceLoss(_LargeMarginSoftmaxLoss):
    """PML 1.6.3 ``CosFaceLoss``: ``logits = s (cos - m onehot)``, cross entropy per element, mean."""

    def __init__(self, *args, margin=0.35, scale=64, **kwargs):
        super().__init__(*args, margin=margin, scale=scale, **kwargs)

    def modify_cosine_of_target_classes(self, cosine_of_target_classes):
        return cosine_of_target_classes - self.margin

    def scale_logits(self, logits, *_):
        return logits * self.scale

    def compute_loss(self, embeddings, labels, indices_tuple, ref_emb, ref_labels):
        logits = (self.get_cosine(embeddings) - self.margin * self.get_target_mask(embeddings, labels)) * self.scale
        return {"loss": {"losses": F.cross_entropy(logits, labels.long(), reduction="none"), "indices": None,
                         "reduction_type": "element"}}


class _Reducer(torch.nn.Module):
    pass


def install() -> None:
    """Register the shim modules and put the reference on ``sys.path`` (idempotent)."""
    if "pytorch_metric_learning" in sys.modules:
        return
    pml = _module("pytorch_metric_learning")
    dist = _module("pytorch_metric_learning.distances")
    dist.BaseDistance = _BaseDistance
    dist.DotProductSimilarity = _DotProductSimilarity
    miners = _module("pytorch_metric_learning.miners")
    miners.BaseMiner = _BaseMiner
    miners.TripletMarginMiner = _TripletMarginMiner
    losses = _module("pytorch_metric_learning.losses")
    losses.BaseMetricLossFunction = _BaseMetricLossFunction
    losses.TripletMarginLoss = _BaseMetricLossFunction
    losses.LargeMarginSoftmaxLoss = _LargeMarginSoftmaxLoss
    losses.CosFaceLoss = _CosFaceLoss
    base = _module("pytorch_metric_learning.losses.base_metric_loss_function")
    base.BaseMetricLossFunction = _BaseMetricLossFunction
    reducers = _module("pytorch_metric_learning.reducers")
    reducers.AvgNonZeroReducer = _Reducer
    utils = _module("pytorch_metric_learning.utils")
    c_f = _module("pytorch_metric_learning.utils.common_functions")
    c_f.to_device = lambda x, tensor=None, device=None, dtype=None: x.to(
        device=device if device is not None else tensor.device, dtype=dtype if dtype is not None else x.dtype)
    c_f.labels_required = lambda labels: None
    c_f.labels_or_indices_tuple_required = lambda labels, indices_tuple: None
    c_f.ref_not_supported = lambda *args: None
    c_f.torch_arange_from_size = lambda x, size_dim=0: torch.arange(x.shape[size_dim], device=x.device)
    lmu = _module("pytorch_metric_learning.utils.loss_and_miner_utils")
    lmu.convert_to_weights = lambda indices_tuple, labels, dtype: torch.ones(labels.shape[0], dtype=dtype, device=labels.device)
    lmu.convert_to_triplets = lambda indices_tuple, labels, ref_labels=None, t_per_anchor=100: indices_tuple
    utils.common_functions = c_f
    utils.loss_and_miner_utils = lmu
    pml.distances, pml.miners, pml.losses, pml.reducers, pml.utils = dist, miners, losses, reducers, utils
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_by_path(name: str, relpath: str):
    """Import one reference file by path (skips package ``__init__`` files that need geoopt)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, f"{REFERENCE_ROOT}/{relpath}")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------
# Whole-model import (tests/test_patch_host.py, oracle/make_golden_layers.py): everything the
# reference's ``hpcs.models`` / ``train.py`` import chain needs besides pytorch-metric-learning.
# Permissive placeholders for plotting / logging / dataset libraries (never called), restated
# behaviour for the few third-party functions the model's forward does call.
# ------------------------------------------------------------------------------------------------
class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything(metaclass=_AnyMeta):
    """Subclassable, callable, attribute-tolerant placeholder."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _PermissiveModule(types.ModuleType):
    """Module whose every missing attribute is a fresh placeholder class (so ``from m import X`` and
    ``class Y(X)`` work)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = _AnyMeta(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


def _permissive(*names):
    for name in names:
        if name not in sys.modules:
            sys.modules[name] = _PermissiveModule(name)
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, sys.modules[name])


class _LightningModule(torch.nn.Module):
    """pytorch_lightning.LightningModule as far as ``BaseSimilarityHypHC`` uses it at construction and in
    ``forward``: hyper-parameter capture is a no-op, ``log`` discards, ``device`` follows the parameters."""
    current_epoch = 0

    def save_hyperparameters(self, *args, **kwargs):
        pass

    def log(self, *args, **kwargs):
        pass

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")


class _Metric(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        return torch.zeros(())


def _random_rotations(n, dtype=None, device=None):
    """pytorch3d 0.7.2 (hpcs-env.yaml:287) ``transforms.random_rotations``: ``randn(n, 4)`` -> the oracle's restatement
    of the quaternion normalisation and the quaternion -> matrix formula."""
    from . import hpcs_oracle
    return hpcs_oracle.quaternion_rotations(torch.randn((n, 4), dtype=dtype, device=device))


class _Rotate:
    """``Rotate(R).transform_points(p)``: row-vector convention, ``p @ R`` per batch element."""

    def __init__(self, R, **kwargs):
        self.R = R

    def transform_points(self, points):
        return torch.bmm(points, self.R.to(points.dtype))


class _RotateAxisAngle(_Rotate):
    """``RotateAxisAngle(angle, axis='Z', degrees=True)`` (the only form the reference uses)."""

    def __init__(self, angle, axis="X", degrees=True, **kwargs):
        from . import hpcs_oracle
        assert axis.upper() == "Z" and degrees
        super().__init__(hpcs_oracle.z_rotations(angle / 360))


def install_models() -> None:
    """``install()`` plus placeholders for pytorch_lightning, torchmetrics, pytorch3d, geoopt, the plotting stack
    and h5py, so that ``import hpcs.models`` and ``import train`` work unmodified on CPU (idempotent)."""
    install()
    if "pytorch_lightning" in sys.modules and hasattr(sys.modules["pytorch_lightning"], "LightningModule"):
        return
    _permissive("pytorch_lightning", "pytorch_lightning.loggers", "pytorch_lightning.callbacks",
                "torchmetrics", "torchmetrics.classification",
                "pytorch3d", "pytorch3d.transforms",
                "geoopt", "geoopt.manifolds", "geoopt.manifolds.stereographic", "geoopt.manifolds.stereographic.math",
                "pyvista", "pyvistaqt", "umap", "h5py",
                "matplotlib", "matplotlib.colors", "matplotlib.pyplot", "matplotlib.collections",
                "mpl_toolkits", "mpl_toolkits.axes_grid1")
    sys.modules["pytorch_lightning"].LightningModule = _LightningModule
    sys.modules["torchmetrics"].Accuracy = _Metric
    sys.modules["torchmetrics.classification"].MulticlassJaccardIndex = _Metric
    t3d = sys.modules["pytorch3d.transforms"]
    t3d.random_rotations, t3d.Rotate, t3d.RotateAxisAngle = _random_rotations, _Rotate, _RotateAxisAngle
