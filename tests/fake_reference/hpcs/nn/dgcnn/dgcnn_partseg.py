import torch
from hpcs.nn.dgcnn.utils.dgcnn_util import get_graph_feature


class DGCNN_partseg(torch.nn.Module):
    pass
