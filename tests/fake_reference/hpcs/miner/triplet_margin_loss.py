import torch
from hpcs import ReferencePathReached


class TripletMarginLoss(torch.nn.Module):
    def __init__(self, margin=0.05, distance=None, **kwargs):
        super().__init__()
        self.margin, self.distance, self.swap, self.smooth_loss = margin, distance, False, False

    def zero_losses(self):
        return {"loss": {"losses": 0, "indices": None, "reduction_type": "already_reduced"}}

    def forward(self, embeddings, labels=None, indices_tuple=None):
        loss_dict = self.compute_loss(embeddings, labels, indices_tuple, embeddings, labels)
        losses = loss_dict["loss"]["losses"]
        if torch.is_tensor(losses) and losses.numel() > 0:
            nz = losses[losses > 0]
            return nz.mean() if nz.numel() else losses.sum() * 0
        return embeddings.sum() * 0

    def compute_loss(self, embeddings, labels, indices_tuple, ref_emb, ref_labels):
        raise ReferencePathReached("TripletMarginLoss.compute_loss: the [n,n] similarity matrix")
