import torch
from torch.nn import functional as F

from hpcs import ReferencePathReached
from hpcs.miner.triplet_margin_miner import RandomTripletMarginMiner
from hpcs.miner.triplet_margin_loss import TripletMarginLoss
from hpcs.distances import hyp_lca, CosineSimilarity


class _CosFace(torch.nn.Module):
    """Plain stand-in for the metric term (outside the hot path)."""

    def __init__(self, num_classes, embedding_size, margin=0.35, scale=2):
        super().__init__()
        self.margin, self.scale = margin, scale
        self.W = torch.nn.Parameter(torch.randn(embedding_size, num_classes))

    def logits(self, embeddings, labels):
        cos = F.normalize(embeddings, dim=1) @ F.normalize(self.W, dim=0)
        return (cos - self.margin * F.one_hot(labels.long(), cos.shape[1])) * self.scale

    def forward(self, embeddings, labels):
        return F.cross_entropy(self.logits(embeddings, labels), labels.long())


class MetricHyperbolicLoss(torch.nn.Module):
    def __init__(self, margin=1.0, t_per_anchor=50, fraction=1.2, scale=1e-3, temperature=0.05, anneal_factor=0.5,
                 num_class=4, embedding_size=4, cosface=True, miner=False):
        super().__init__()
        self.margin, self.t_per_anchor, self.fraction, self.scale = margin, t_per_anchor, fraction, scale
        self.temperature, self.anneal_factor = temperature, anneal_factor
        self.num_class, self.embedding_size, self.cosface, self.miner = num_class, embedding_size, cosface, miner
        self.distance_sim = CosineSimilarity()
        if self.miner:
            self.hyp_miner = RandomTripletMarginMiner(distance=self.distance_sim, margin=0, t_per_anchor=t_per_anchor,
                                                      fraction=fraction, type_of_triplets='easy')
        if self.cosface:
            self.loss_cosface = _CosFace(num_class, embedding_size)
        else:
            self.triplet_miner = RandomTripletMarginMiner(distance=self.distance_sim, margin=margin,
                                                          t_per_anchor=t_per_anchor, fraction=fraction,
                                                          type_of_triplets='semihard')
            self.loss_triplet = TripletMarginLoss(distance=self.distance_sim, margin=margin)

    def get_triplets(self, n_samples):
        ij = torch.combinations(torch.arange(n_samples), r=2).repeat_interleave(self.t_per_anchor, dim=0)
        k = torch.randint(n_samples, (ij.shape[0],), dtype=torch.long)
        ok = (ij[:, 0] != k) & (ij[:, 1] != k)
        return ij[ok, 0], ij[ok, 1], k[ok]

    def compute_hyp(self, x_poincare, labels):
        raise ReferencePathReached("MetricHyperbolicLoss.compute_hyp")

    def get_logits(self, embeddings, labels):
        return self.loss_cosface.logits(embeddings, labels)

    def compute_loss(self, x_euclidean, x_poincare, labels, *args):
        loss_hyperbolic = self.compute_hyp(x_poincare, labels)
        if self.cosface:
            loss_metric = self.loss_cosface(x_poincare, labels.long())
        else:
            loss_metric = self.loss_triplet(x_poincare, labels, self.triplet_miner(x_poincare, labels))
        return {"loss_hyp": {"losses": loss_hyperbolic}, "loss_metric": {"losses": loss_metric}}

    def normalize_embeddings(self, embeddings):
        return F.normalize(embeddings, p=2, dim=1) * torch.clamp(self.scale, 1e-4, 1)


class HierarchicalMetricHyperbolicLoss(MetricHyperbolicLoss):
    """Subclass bound at import time to the class above -- the PartNet default."""

    def __init__(self, hierarchy_list=(), **kwargs):
        super().__init__(cosface=True, **kwargs)
        self.hierarchy_list = hierarchy_list

    def compute_loss(self, x_euclidean, x_poincare, labels, *args):
        return {"loss_hyp": {"losses": self.compute_hyp(x_poincare, labels)},
                "loss_metric": {"losses": self.loss_cosface(x_poincare, labels.long())}}
