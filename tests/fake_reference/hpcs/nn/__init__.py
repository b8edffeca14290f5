import torch


def MLP(channels, bias=True, negative_slope=0.0, dropout=0.0):
    return torch.nn.Sequential(*[torch.nn.Linear(channels[i - 1], channels[i], bias=bias) for i in range(1, len(channels))])
