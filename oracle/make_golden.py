"""Generate ``tests/golden/*.npz`` by executing the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Run once here (``python -m oracle.make_golden``); the outputs are small
and committed, because ``/root/reference`` does not exist on the GPU box.  Every array stored is
either a seeded synthetic input or an output of the reference's own functions:

  knn / get_graph_feature(_cross) : hpcs/nn/dgcnn/utils/vn_dgcnn_util.py (cross: the pointnet copy,
                                    whose device pick works on CPU)
  hyp_lca                         : hpcs/distances/lca.py
  expmap_1 / project              : hpcs/utils/poincare.py, hpcs/distances/poincare.py
  get_balanced_random_triplet_indices, RandomTripletMarginMiner : hpcs/miner/
  MetricHyperbolicLoss.compute_hyp: hpcs/loss/ultrametric_loss.py  (fp32 and fp64 evaluations)
  _decode_linkage                 : the three statements of hpcs/models/base_hyp_hc.py:83-85 using
                                    the reference's normalize_embeddings/project + scipy linkage
                                    (the LightningModule itself needs pytorch_lightning/pytorch3d).

pytorch-metric-learning is replaced by the shims in ``oracle/ref_stubs.py``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_stubs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_cloud(gen, B, N):
    """N(0,1) points, centred and scaled to unit max radius (hpcs/utils/data.py:37-42)."""
    pts = torch.randn(B, N, 3, generator=gen)
    pts = pts - pts.mean(dim=1, keepdim=True)
    pts = pts / pts.norm(dim=-1).amax(dim=1).view(B, 1, 1)
    return pts.transpose(1, 2).contiguous()          # [B,3,N]


def synth_labels(gen, B, N, n_parts=(2, 6), first_id=0):
    labs = []
    for b in range(B):
        parts = int(torch.randint(n_parts[0], n_parts[1] + 1, (1,), generator=gen))
        probs = torch._sample_dirichlet(torch.ones(parts), generator=gen)      # seeded: the fixture regenerates bit for bit
        ids = torch.multinomial(probs, N, replacement=True, generator=gen) + first_id
        first_id += parts
        labs.append(ids)
    return torch.cat(labs)


def main():
    ref_stubs.install()
    torch.set_num_threads(1)           # fp32 index_put accumulation order: one thread -> bit-reproducible fixtures
    os.makedirs(OUT, exist_ok=True)
    util = ref_stubs.load_by_path("ref_vn_dgcnn_util", "hpcs/nn/dgcnn/utils/vn_dgcnn_util.py")
    util_pn = ref_stubs.load_by_path("ref_vn_dgcnn_util_pn", "hpcs/nn/pointnet/utils/vn_dgcnn_util.py")
    from hpcs.distances import hyp_lca
    from hpcs.distances.poincare import project
    from hpcs.utils.poincare import expmap_1
    from hpcs.miner.loss_and_miner_utils import get_balanced_random_triplet_indices
    from hpcs.loss.ultrametric_loss import MetricHyperbolicLoss
    from scipy.cluster.hierarchy import linkage

    gen = torch.Generator().manual_seed(0)

    # ---- kNN -------------------------------------------------------------------------------
    x3 = synth_cloud(gen, 2, 256)
    x63 = torch.randn(2, 63, 128, generator=gen)
    np.savez_compressed(os.path.join(OUT, "knn.npz"),
                        x3=x3.numpy(), idx3=util.knn(x3, 20).numpy().astype(np.int32),
                        x63=x63.numpy(), idx63=util.knn(x63, 10).numpy().astype(np.int32))

    # ---- edge features (fwd + bwd) -----------------------------------------------------------
    xg = torch.randn(2, 3, 3, 48, generator=gen, requires_grad=True)
    out = util.get_graph_feature(xg, k=6)
    gout = torch.randn(out.shape, generator=gen)
    (gx,) = torch.autograd.grad(out, xg, gout)
    idx_g = util.knn(xg.detach().view(2, 9, 48), 6)
    xc = torch.randn(2, 1, 3, 40, generator=gen, requires_grad=True)
    idx_c = util.knn(xc.detach().view(2, 3, 40), 5)
    outc = util_pn.get_graph_feature_cross(xc, k=5, idx=idx_c.clone())
    goutc = torch.randn(outc.shape, generator=gen)
    (gxc,) = torch.autograd.grad(outc, xc, goutc)
    # fixed graph from coordinates (x_coord=...) as in the non-dynamic mode
    coord = synth_cloud(gen, 2, 48)
    out_fixed = util.get_graph_feature(xg.detach(), k=6, x_coord=coord)
    np.savez_compressed(os.path.join(OUT, "edge_feat.npz"),
                        x=xg.detach().numpy(), idx=idx_g.numpy().astype(np.int32), out=out.detach().numpy(),
                        gout=gout.numpy(), gx=gx.numpy(),
                        xc=xc.detach().numpy(), idxc=idx_c.numpy().astype(np.int32),
                        outc=outc.detach().numpy(), goutc=goutc.numpy(), gxc=gxc.numpy(),
                        coord=coord.numpy(), out_fixed=out_fixed.numpy())

    # ---- hyp_lca (fp64 evaluation of the reference; fp32 reference beside it) -------------------
    rec = {}
    for tag, radius in (("s1e-3", 1e-3), ("s1e-2", 1e-2), ("s0.1", 0.1), ("s0.5", 0.5), ("s0.9", 0.9), ("mixed", None)):
        a = torch.randn(48, 32, generator=gen)
        b = torch.randn(48, 32, generator=gen)
        if radius is None:      # unequal norms, general signature
            a = a / a.norm(dim=-1, keepdim=True) * (0.05 + 0.85 * torch.rand(48, 1, generator=gen))
            b = b / b.norm(dim=-1, keepdim=True) * (0.05 + 0.85 * torch.rand(48, 1, generator=gen))
        else:
            a = a / a.norm(dim=-1, keepdim=True) * radius
            b = b / b.norm(dim=-1, keepdim=True) * radius
        ad = a.double().requires_grad_(True)
        bd = b.double().requires_grad_(True)
        dist = hyp_lca(ad, bd, return_coord=False)
        ga, gb = torch.autograd.grad(dist.sum(), (ad, bd))
        coordp = hyp_lca(ad, bd, return_coord=True)
        gc = torch.randn(48, 32, generator=gen).double()
        gca, gcb = torch.autograd.grad((coordp * gc).sum(), (ad, bd))
        rec.update({f"{tag}_a": a.numpy(), f"{tag}_b": b.numpy(), f"{tag}_dist": dist.detach().numpy(),
                    f"{tag}_ga": ga.numpy(), f"{tag}_gb": gb.numpy(), f"{tag}_coord": coordp.detach().numpy(),
                    f"{tag}_gc": gc.numpy(), f"{tag}_gca": gca.numpy(), f"{tag}_gcb": gcb.numpy(),
                    f"{tag}_dist_ref32": hyp_lca(a, b, return_coord=False).numpy()})
    np.savez_compressed(os.path.join(OUT, "hyp_lca.npz"), **rec)

    # ---- expmap / project --------------------------------------------------------------------
    u = torch.randn(64, 32, generator=gen) * torch.logspace(-3, 1.5, 64).view(64, 1)
    ball = expmap_1(u, torch.zeros_like(u))
    np.savez_compressed(os.path.join(OUT, "expmap.npz"), u=u.numpy(), y=ball.numpy(),
                        y64=expmap_1(u.double(), torch.zeros_like(u.double())).numpy(),
                        proj=project(ball * 1.001).numpy())

    # ---- sampler + miner + compute_hyp -------------------------------------------------------
    n, D = 384, 32
    labels = synth_labels(gen, 3, 128)
    x = expmap_1(torch.randn(n, D, generator=gen), torch.zeros(n, D))
    rec = {"x": x.numpy(), "labels": labels.numpy().astype(np.int32)}
    for frac in (0.0, 1.2):
        torch.manual_seed(1234)
        a, p, ng = get_balanced_random_triplet_indices(labels, t_per_anchor=7, fraction=frac)
        rec.update({f"f{frac}_a": a.numpy().astype(np.int32), f"f{frac}_p": p.numpy().astype(np.int32),
                    f"f{frac}_n": ng.numpy().astype(np.int32)})
    for tag, scale, temp in (("s1e-3", 1e-3, 0.05), ("s0.1", 0.1, 0.05), ("s0.5", 0.5, 0.1)):
        for dt, name in ((torch.float64, "64"), (torch.float32, "32")):
            sc = torch.nn.Parameter(torch.tensor([scale], dtype=dt))
            loss_mod = MetricHyperbolicLoss(margin=0.35, t_per_anchor=7, fraction=0.0, scale=sc,
                                            temperature=temp, num_class=50, embedding_size=D,
                                            cosface=True, miner=True)
            xd = x.to(dt).requires_grad_(True)
            torch.manual_seed(1234)
            kept = loss_mod.hyp_miner(xd.detach(), labels)
            torch.manual_seed(1234)
            loss = loss_mod.compute_hyp(xd, labels)
            gx_, gs_ = torch.autograd.grad(loss, (xd, sc))
            rec.update({f"{tag}_loss{name}": loss.detach().numpy(), f"{tag}_gx{name}": gx_.numpy(),
                        f"{tag}_gscale{name}": gs_.numpy(), f"{tag}_kept{name}": np.int64(kept[0].numel()),
                        f"{tag}_kept_a{name}": kept[0].numpy().astype(np.int32)})
        rec[f"{tag}_scale"] = np.float32(scale)
        rec[f"{tag}_temp"] = np.float32(temp)
    np.savez_compressed(os.path.join(OUT, "compute_hyp.npz"), **rec)

    # ---- decode ------------------------------------------------------------------------------
    rec = {}
    sc = torch.nn.Parameter(torch.tensor([1e-3]))
    loss_mod = MetricHyperbolicLoss(scale=sc, num_class=4, embedding_size=32, miner=True)
    for N in (96, 200):
        e = expmap_1(torch.randn(N, 32, generator=gen), torch.zeros(N, 32))
        leaves = project(loss_mod.normalize_embeddings(e)).detach().cpu()      # base_hyp_hc.py:83-84
        rec[f"x{N}"] = e.numpy()
        rec[f"Zc{N}"] = linkage(leaves, method="complete", metric="cosine")   # base_hyp_hc.py:85
        rec[f"Zs{N}"] = linkage(leaves, method="single", metric="cosine")
    # structured (clustered) embeddings: 5 tight clusters, as after training
    cen = torch.randn(5, 32, generator=gen)
    e = expmap_1(cen[torch.randint(0, 5, (150,), generator=gen)] + 0.05 * torch.randn(150, 32, generator=gen),
                 torch.zeros(150, 32))
    leaves = project(loss_mod.normalize_embeddings(e)).detach().cpu()
    rec["xclu"] = e.numpy()
    rec["Zcclu"] = linkage(leaves, method="complete", metric="cosine")
    rec["Zsclu"] = linkage(leaves, method="single", metric="cosine")
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **rec)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
