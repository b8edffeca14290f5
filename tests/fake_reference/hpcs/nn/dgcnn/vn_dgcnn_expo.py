import torch
from hpcs.nn.dgcnn.utils.vn_dgcnn_util import get_graph_feature


class VN_DGCNN_expo(torch.nn.Module):
    pass
