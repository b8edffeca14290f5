import torch
from hpcs.nn.dgcnn.utils.vn_dgcnn_util import get_graph_feature


class VN_DGCNN_partseg(torch.nn.Module):
    """Backbone-SHAPED stand-in: the same three graph-feature calls as the reference's forward (a D=3 graph, then two
    D=63 graphs on 21 vector channels), each followed by a channel-mixing linear map and a mean over the k neighbours,
    then rotation-invariant norms -> per-point features.  Not the reference's network."""

    def __init__(self, in_channels, out_features, k, dropout, pooling, num_categories):
        super().__init__()
        self.k, self.out_features, self.num_categories = k, out_features, num_categories
        self.mix1 = torch.nn.Linear(2, 21, bias=False)
        self.mix2 = torch.nn.Linear(42, 21, bias=False)
        self.mix3 = torch.nn.Linear(42, 21, bias=False)
        self.head = torch.nn.Linear(63 + num_categories, out_features)

    def _layer(self, x, mix):
        e = get_graph_feature(x, k=self.k)                      # [B,2C,3,N,k]
        return mix(e.transpose(1, -1)).transpose(1, -1).mean(dim=-1)

    def forward(self, x, l):
        x1 = self._layer(x.unsqueeze(1), self.mix1)
        x2 = self._layer(x1, self.mix2)
        x3 = self._layer(x2, self.mix3)
        inv = torch.cat((x1, x2, x3), dim=1).norm(dim=2)        # [B,63,N]
        cat = l.reshape(l.shape[0], -1, 1).expand(-1, -1, inv.shape[-1])
        return self.head(torch.cat((inv, cat), dim=1).transpose(1, 2))
