import torch
from hpcs import ReferencePathReached


class CosineSimilarity(torch.nn.Module):
    is_inverted = True

    def forward(self, query_emb, ref_emb=None):
        raise ReferencePathReached("CosineSimilarity: the [n,n] similarity matrix")

    def pairwise_distance(self, query_emb, ref_emb):
        return 0.5 * (1 + torch.sum(query_emb * ref_emb, dim=1))

    def margin(self, x, y):
        return y - x

    def smallest_dist(self, *args, **kwargs):
        return torch.max(*args, **kwargs)
