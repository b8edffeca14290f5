import torch
from hpcs import ReferencePathReached
from hpcs.miner.loss_and_miner_utils import get_balanced_random_triplet_indices


class RandomTripletMarginMiner(torch.nn.Module):
    def __init__(self, t_per_anchor, fraction, margin=0.2, type_of_triplets="all", distance=None, **kwargs):
        super().__init__()
        self.t_per_anchor, self.fraction = t_per_anchor, fraction
        self.margin, self.type_of_triplets, self.distance = margin, type_of_triplets, distance

    def forward(self, embeddings, labels, ref_emb=None, ref_labels=None):
        with torch.no_grad():
            return self.mine(embeddings, labels, embeddings, labels)

    def mine(self, embeddings, labels, ref_emb, ref_labels):
        raise ReferencePathReached("RandomTripletMarginMiner.mine")
