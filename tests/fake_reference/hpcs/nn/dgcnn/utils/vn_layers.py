"""Stand-ins with the reference's class and attribute names (hpcs/nn/dgcnn/utils/vn_layers.py); simplified bodies."""
import torch

from hpcs import ReferencePathReached


class VNBatchNorm(torch.nn.Module):
    def __init__(self, num_features, dim):
        super().__init__()
        self.bn = torch.nn.BatchNorm2d(num_features) if dim == 5 else torch.nn.BatchNorm1d(num_features)


EPS = 1e-6


class VNLinearLeakyReLU(torch.nn.Module):
    """Parameters where the reference keeps them (map_to_feat, map_to_dir, batchnorm.bn).  The 5-D (edge tensor) form is a
    hot-path consumer: calling it means the fused layer was not bound.  The 4-D form (conv6 / std-feature, outside the
    path) does the layer's arithmetic with library ops -- Linear over channels, BatchNorm on the vector norms, leaky
    projection onto the learnt direction -- so the dense tail of the backbone runs at its real cost."""

    def __init__(self, in_channels, out_channels, dim=5, share_nonlinearity=False, negative_slope=0.2):
        super().__init__()
        self.dim, self.negative_slope = dim, negative_slope
        self.map_to_feat = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.map_to_dir = torch.nn.Linear(in_channels, 1 if share_nonlinearity else out_channels, bias=False)
        self.batchnorm = VNBatchNorm(out_channels, dim=dim)

    def forward(self, x):
        if x.dim() == 5:
            raise ReferencePathReached("VNLinearLeakyReLU on the [B,2C,3,N,k] edge tensor")
        p = self.map_to_feat(x.transpose(1, -1)).transpose(1, -1)
        norm = p.norm(dim=2) + EPS
        p = p / norm.unsqueeze(2) * self.batchnorm.bn(norm).unsqueeze(2)
        d = self.map_to_dir(x.transpose(1, -1)).transpose(1, -1)
        dot = (p * d).sum(2, keepdim=True)
        keep = (dot >= 0).to(p.dtype)
        proj = p - (dot / ((d * d).sum(2, keepdim=True) + EPS)) * d
        s = self.negative_slope
        return s * p + (1 - s) * (keep * p + (1 - keep) * proj)


class VNStdFeature(torch.nn.Module):
    def __init__(self, in_channels, dim=4, normalize_frame=False, share_nonlinearity=False, negative_slope=0.2):
        super().__init__()
        self.vn1 = VNLinearLeakyReLU(in_channels, in_channels // 2, dim=dim, share_nonlinearity=share_nonlinearity, negative_slope=negative_slope)
        self.vn2 = VNLinearLeakyReLU(in_channels // 2, in_channels // 4, dim=dim, share_nonlinearity=share_nonlinearity, negative_slope=negative_slope)
        self.vn_lin = torch.nn.Linear(in_channels // 4, 3, bias=False)

    def forward(self, x):
        z0 = self.vn2(self.vn1(x))
        z0 = self.vn_lin(z0.transpose(1, -1)).transpose(1, -1).transpose(1, 2)       # [B,3,3,N]
        return torch.einsum('bijm,bjkm->bikm', x, z0), z0


def mean_pool(x, dim=-1, keepdim=False):
    if x.dim() == 5:
        raise ReferencePathReached("mean_pool on the edge tensor")
    return x.mean(dim=dim, keepdim=keepdim)
