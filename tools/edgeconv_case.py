"""Diagnose one fused-EdgeConv case against the fp64 oracle: errors of every output, and how concentrated the input-gradient
error is (a direction-test flip of one (edge, channel) pair touches a handful of points).
  python tools/edgeconv_case.py B C N k two train"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpcs_b200 as hb  # noqa: E402
import test_gpu_edgeconv as T  # noqa: E402
from oracle import hpcs_oracle as O  # noqa: E402


def main():
    B, C, N, k = (int(v) for v in sys.argv[1:5])
    two, train = sys.argv[5] == "1", sys.argv[6] == "1"
    from hpcs_b200.edgeconv import edgeconv
    gen = torch.Generator().manual_seed(B * 1000 + N + k)
    torch.manual_seed(N + k)
    convs = [T.VNConv(2 * C)] + ([T.VNConv(21)] if two else [])
    for c in convs:
        bn = c.batchnorm.bn
        with torch.no_grad():
            bn.weight.copy_(0.5 + torch.rand(21, generator=gen)); bn.bias.copy_(0.3 * torch.randn(21, generator=gen))
            bn.running_mean.copy_(1.0 + torch.rand(21, generator=gen)); bn.running_var.copy_(0.5 + torch.rand(21, generator=gen))
        c.cuda().train(train)
    x = torch.randn(B, C, 3, N, generator=gen)
    gout = torch.randn(B, 21, 3, N, generator=gen)
    xg = x.cuda().requires_grad_(True)
    idx = hb.knn(xg.detach().view(B, 3 * C, N), k)
    oc = [c.as_oracle() for c in convs]
    y = edgeconv(xg, k, convs[0], convs[1] if two else None, idx=idx)
    grads = torch.autograd.grad((y * gout.cuda()).sum(), [xg] + T.param_list(convs))
    xd = x.double().requires_grad_(True)
    yw = O.edgeconv_layer(xd, idx.cpu(), oc, training=train)
    wparams = [c[n] for c in oc for n in ("wf", "wd", "gamma", "beta")]
    wgrads = torch.autograd.grad((yw * gout.double()).sum(), [xd] + wparams)
    print("forward", T.nrm_err(y, yw))
    dy = (y.detach().cpu().double() - yw.detach()).abs()
    print("  forward max abs err", dy.max().item(), "at", [int(v) for v in torch.nonzero(dy == dy.max())[0]], "ref magnitude", yw.abs().mean().item())
    names = ["x"] + [f"conv{j}.{n}" for j in range(len(convs)) for n in ("wf", "wd", "gamma", "beta")]
    for nme, g, w in zip(names, grads, wgrads):
        print(f"grad {nme:12s} {T.nrm_err(g, w):.3e}")
    d = (grads[0].detach().cpu().double() - wgrads[0]).pow(2).sum(dim=(1, 2))            # [B, N] squared error per point
    tot = d.sum().item()
    top = torch.topk(d.reshape(-1), 8)
    print("share of the squared input-gradient error in the 8 worst points:", (top.values.sum().item() / tot), "of", d.numel(), "points")
    if train:
        for j, (c, o) in enumerate(zip(convs, oc)):
            print(f"running_mean conv{j}", (c.batchnorm.bn.running_mean.cpu().double() - o["running_mean"]).abs().max().item(),
                  "running_var", (c.batchnorm.bn.running_var.cpu().double() - o["running_var"]).abs().max().item(), "/", o["running_var"].abs().max().item())


if __name__ == "__main__":
    main()
