import torch

from hpcs import ReferencePathReached
from hpcs.loss.ultrametric_loss import MetricHyperbolicLoss
from hpcs.utils.scores import get_optimal_k
from hpcs.utils.data import to_categorical


class BaseSimilarityHypHC(torch.nn.Module):
    """Calling structure of the reference's LightningModule (forward -> _forward -> losses [-> decode])."""

    def __init__(self, nn_feat, nn_emb, euclidean_size, hyp_size, margin=0.5, t_per_anchor=50, fraction=1.2,
                 temperature=0.05, anneal_factor=0.5, num_class=4, trade_off=0.1, miner=True, cosface=True, **kwargs):
        super().__init__()
        self.nn_feat, self.nn_emb = nn_feat, nn_emb
        self.euclidean_size, self.hyp_size, self.num_class, self.trade_off = euclidean_size, hyp_size, num_class, trade_off
        self.margin, self.t_per_anchor, self.fraction = margin, t_per_anchor, fraction
        self.temperature, self.anneal_factor, self.miner, self.cosface = temperature, anneal_factor, miner, cosface
        self.scale = torch.nn.Parameter(torch.Tensor([1e-3]), requires_grad=True)
        self.metric_hyp_loss = MetricHyperbolicLoss(margin=margin, t_per_anchor=t_per_anchor, fraction=fraction,
                                                    scale=self.scale, temperature=temperature,
                                                    anneal_factor=anneal_factor, num_class=num_class,
                                                    embedding_size=hyp_size, miner=miner, cosface=cosface)

    @property
    def device(self):
        return self.scale.device

    def _decode_linkage(self, leaves_embeddings):
        raise ReferencePathReached("BaseSimilarityHypHC._decode_linkage (scipy on the host)")

    def _forward(self, batch, testing):
        raise NotImplementedError

    def compute_losses(self, x_euclidean, x_poincare, labels):
        loss = self.metric_hyp_loss.compute_loss(x_euclidean, x_poincare, labels.view(-1, 1)[:, 0].long())
        return {"loss_metric": loss["loss_metric"]["losses"], "loss_hyp": loss["loss_hyp"]["losses"] * self.trade_off}

    def compute_accuracy(self, embeddings, labels):
        return (self.metric_hyp_loss.get_logits(embeddings, labels).argmax(1) == labels).float().mean()

    def compute_iou(self, embeddings, labels):
        return self.compute_accuracy(embeddings, labels)

    def forward(self, batch, testing=False):
        """The reference's calling structure, per-cloud decode loop included (replaced by the binding)."""
        points, x_euclidean, x_poincare, pts_labels = self._forward(batch, testing)
        xe = x_euclidean.contiguous().view(-1, x_euclidean.shape[-1])
        xp = x_poincare.contiguous().view(-1, x_poincare.shape[-1])
        losses = self.compute_losses(xe, xp, pts_labels)
        if not testing:
            return losses, {}
        Z = [self._decode_linkage(x_poincare[i]) for i in range(points.size(0))]
        return losses, {}, x_euclidean, x_poincare, Z, points, pts_labels

    def test_scores(self, targets, linkage_matrix):
        return [get_optimal_k(targets[i].cpu(), linkage_matrix[i], 'iou') for i in range(len(linkage_matrix))]
