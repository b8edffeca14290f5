import torch
from hpcs.nn.pointnet.utils.vn_dgcnn_util import get_graph_feature_cross
from hpcs.nn.pointnet import vn_pointnet  # noqa: F401  (holds its own alias, like the reference)


class VN_POINTNET_partseg(torch.nn.Module):
    def __init__(self, k=20):
        super().__init__()
        self.k = k

    def forward(self, x):
        return get_graph_feature_cross(x.unsqueeze(1), k=self.k).mean(dim=-1)
