import torch
import torch.nn as nn
from hpcs.nn.dgcnn.utils.vn_layers import *
from hpcs.nn.dgcnn.utils.vn_dgcnn_util import get_graph_feature


class VN_DGCNN_partseg(nn.Module):
    """The reference's constructor signature, attribute names and layer shapes for the graph layers (conv1 .. conv5 on
    64 // 3 = 21 vector channels); a narrower dense tail by default (``tail_width`` = 32 keeps the tests small; set the class
    attribute to 1024 // 3 = 341 and every layer has the reference's shape: conv8 then takes 2299 channels and the module
    holds the reference's 1.3 M parameters -- what ``bench.py``'s ``model_step`` record uses).  ``forward`` is the
    reference's call sequence for the first layer only -- enough to reach ``get_graph_feature`` -- because every later
    statement is replaced by the binding."""
    tail_width = 32

    def __init__(self, in_channels, out_features, k, dropout, pooling, num_categories):
        super().__init__()
        self.in_channels, self.out_features, self.k = in_channels, out_features, k
        self.dropout, self.pooling, self.num_categories = dropout, pooling, num_categories
        self.conv1 = VNLinearLeakyReLU(2, 64 // 3)
        self.conv2 = VNLinearLeakyReLU(64 // 3, 64 // 3)
        self.conv3 = VNLinearLeakyReLU(64 // 3 * 2, 64 // 3)
        self.conv4 = VNLinearLeakyReLU(64 // 3, 64 // 3)
        self.conv5 = VNLinearLeakyReLU(64 // 3 * 2, 64 // 3)
        self.pool1 = self.pool2 = self.pool3 = mean_pool
        wide = self.tail_width
        self.conv6 = VNLinearLeakyReLU(64 // 3 * 3, wide, dim=4, share_nonlinearity=True)
        self.std_feature = VNStdFeature(wide * 2, dim=4, normalize_frame=False)
        self.conv7 = nn.Sequential(nn.Conv1d(num_categories, 64, kernel_size=1, bias=False), nn.BatchNorm1d(64), nn.LeakyReLU(0.2))
        self.conv8 = nn.Sequential(nn.Conv1d(wide * 2 * 3 + 64 + 63 * 3, 256, kernel_size=1, bias=False), nn.BatchNorm1d(256), nn.LeakyReLU(0.2))
        self.dp1 = nn.Dropout(p=dropout)
        self.conv9 = nn.Sequential(nn.Conv1d(256, 256, kernel_size=1, bias=False), nn.BatchNorm1d(256), nn.LeakyReLU(0.2))
        self.dp2 = nn.Dropout(p=dropout)
        self.conv10 = nn.Sequential(nn.Conv1d(256, 128, kernel_size=1, bias=False), nn.BatchNorm1d(128), nn.LeakyReLU(0.2))
        self.conv11 = nn.Sequential(nn.Conv1d(128, out_features, kernel_size=1, bias=False), nn.BatchNorm1d(out_features))

    def forward(self, x, l):
        x = get_graph_feature(x.unsqueeze(1), k=self.k)
        return self.pool1(self.conv2(self.conv1(x)))
