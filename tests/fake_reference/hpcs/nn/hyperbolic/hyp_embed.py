import torch
from hpcs import ReferencePathReached
from hpcs.nn import MLP


class ExpMap(torch.nn.Module):
    def forward(self, x):
        raise ReferencePathReached("ExpMap.forward")


class MLPExpMap(torch.nn.Module):
    def __init__(self, input_feat, out_feat, bias=False, negative_slope=0.2, dropout=0.0):
        super().__init__()
        self.mlp = MLP([input_feat, out_feat], bias=bias)

    def forward(self, x):
        raise ReferencePathReached("MLPExpMap.forward")
