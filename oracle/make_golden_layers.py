"""Generate ``tests/golden/edgeconv.npz`` by executing the UNMODIFIED reference layers in this container.

TEST INFRASTRUCTURE ONLY (``python -m oracle.make_golden_layers``).  For three layer shapes of
``VN_DGCNN_partseg.forward`` (hpcs/nn/dgcnn/vn_dgcnn_partseg.py:65-77) -- conv1+conv2 on the coordinates (C=1), conv3+conv4
on 21 vector channels, conv5 alone -- it runs the reference's own ``get_graph_feature`` (vn_dgcnn_util.py),
``VNLinearLeakyReLU`` modules and ``mean_pool`` (vn_layers.py) in training and in eval mode and stores inputs, the kNN
graph used, every parameter and buffer, the layer output, the gradients of a seeded linear functional wrt the input and
every parameter, and the BatchNorm buffers after the training-mode call.  Each case is evaluated twice by the reference
code: in fp32 (what it ships; keys ``*_y``, ``*_gx`` ...) and in fp64 (``*_y64`` ...: the parity target -- for the C=1 layer the
reference's own fp32 gradients sit up to 2e-2 from its fp64 ones, because a vector channel whose norm is nearly 0 on some
edge has a 1/|p| gradient).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_stubs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.set_num_threads(1)
    util = ref_stubs.load_by_path("ref_vn_dgcnn_util", "hpcs/nn/dgcnn/utils/vn_dgcnn_util.py")
    layers = ref_stubs.load_by_path("ref_vn_layers", "hpcs/nn/dgcnn/utils/vn_layers.py")
    gen = torch.Generator().manual_seed(21)
    rec = {}
    B, N, k = 2, 40, 6
    for tag, C, n_conv in (("l1", 1, 2), ("l2", 21, 2), ("l3", 21, 1)):
        torch.manual_seed(100 + C + n_conv)
        convs = [layers.VNLinearLeakyReLU(2 * C, 21)] + ([layers.VNLinearLeakyReLU(21, 21)] if n_conv == 2 else [])
        for c in convs:                                        # non-trivial BatchNorm state (a fresh one is gamma=1, beta=0)
            bn = c.batchnorm.bn
            with torch.no_grad():
                bn.weight.copy_(0.5 + torch.rand(21, generator=gen))
                bn.bias.copy_(0.3 * torch.randn(21, generator=gen))
                bn.running_mean.copy_(0.5 + 0.5 * torch.rand(21, generator=gen))
                bn.running_var.copy_(0.2 + torch.rand(21, generator=gen))
        x = torch.randn(B, C, 3, N, generator=gen)
        idx = util.knn(x.view(B, 3 * C, N), k)
        gout = torch.randn(B, 21, 3, N, generator=gen)
        rec.update({f"{tag}_x": x.numpy(), f"{tag}_idx": idx.numpy().astype(np.int32), f"{tag}_gout": gout.numpy()})
        for j, c in enumerate(convs):
            bn = c.batchnorm.bn
            rec.update({f"{tag}_c{j}_wf": c.map_to_feat.weight.detach().numpy(), f"{tag}_c{j}_wd": c.map_to_dir.weight.detach().numpy(),
                        f"{tag}_c{j}_gamma": bn.weight.detach().numpy(), f"{tag}_c{j}_beta": bn.bias.detach().numpy(),
                        f"{tag}_c{j}_rm": bn.running_mean.numpy().copy(), f"{tag}_c{j}_rv": bn.running_var.numpy().copy()})
        import copy
        convs64 = [copy.deepcopy(c).double() for c in convs]
        for mode in ("eval", "train"):                         # eval first: the training call updates the buffers
            for suffix, mods, dt in (("", convs, torch.float32), ("64", convs64, torch.float64)):
                for c in mods:
                    c.train(mode == "train")
                xr = x.to(dt).clone().requires_grad_(True)
                e = util.get_graph_feature(xr, k=k, idx=idx.clone())
                for c in mods:
                    e = c(e)
                y = layers.mean_pool(e)
                params = [p for c in mods for p in (c.map_to_feat.weight, c.map_to_dir.weight, c.batchnorm.bn.weight, c.batchnorm.bn.bias)]
                grads = torch.autograd.grad((y * gout.to(dt)).sum(), [xr] + params)
                rec[f"{tag}_{mode}_y{suffix}"] = y.detach().numpy()
                rec[f"{tag}_{mode}_gx{suffix}"] = grads[0].numpy()
                for j in range(len(mods)):
                    for q, name in enumerate(("gwf", "gwd", "ggamma", "gbeta")):
                        rec[f"{tag}_{mode}_c{j}_{name}{suffix}"] = grads[1 + 4 * j + q].numpy()
        for j, c in enumerate(convs):
            rec[f"{tag}_c{j}_rm_after"] = c.batchnorm.bn.running_mean.numpy().copy()
            rec[f"{tag}_c{j}_rv_after"] = c.batchnorm.bn.running_var.numpy().copy()
    path = os.path.join(OUT, "edgeconv.npz")
    np.savez_compressed(path, **rec)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
