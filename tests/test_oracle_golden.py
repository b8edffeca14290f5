"""Pin the oracle (oracle/hpcs_oracle.py, oracle/knn_canonical.c) to outputs of the reference itself
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import hpcs_oracle as O


def t(a, dtype=None):
    out = torch.from_numpy(np.asarray(a))
    return out if dtype is None else out.to(dtype)


def ambiguous_rows(x, k, tol_scale=64.0):
    """Rows whose fp64 ranking has a gap below the fp32 error bound between any two adjacent ranks
    among the first k+1 (SURVEY.md Finding 5): index parity is only defined outside these rows."""
    d = O.neg_sqdist_fp64(x)
    top = d.topk(k + 1, dim=-1)[0]
    gaps = top[..., :-1] - top[..., 1:]
    sq = (x.double() ** 2).sum(1)
    bound = tol_scale * np.finfo(np.float32).eps * (sq.unsqueeze(-1) + sq.amax(dim=-1, keepdim=True).unsqueeze(-1))
    return (gaps < bound).any(-1)


@pytest.mark.parametrize("key,k", [("3", 20), ("63", 10)])
def test_knn_reference_and_canonical(golden, key, k):
    g = golden("knn")
    x, ref = t(g["x" + key]), t(g["idx" + key], torch.int64)
    assert torch.equal(O.knn_reference(x, k), ref)               # restatement == reference, same torch
    can = O.knn_canonical(x, k)
    amb = ambiguous_rows(x, k)
    assert torch.equal(can[~amb], ref[~amb])                     # canonical order == reference off near-ties
    assert amb.float().mean() < 0.05
    # near-tie rows: same neighbour SET after an fp64 re-rank
    d = O.neg_sqdist_fp64(x)
    exact = d.topk(k, dim=-1)[1]
    for b, i in amb.nonzero().tolist():
        kth = d[b, i, exact[b, i, -1]]
        assert (d[b, i, can[b, i]] >= kth - 1e-6 * abs(kth) - 1e-12).all()
    # canonical order is sorted by (value desc, index asc) and contains self
    _, val = O.knn_canonical(x, k, return_values=True)
    assert (val[..., :-1] >= val[..., 1:]).all()
    assert (can == torch.arange(x.shape[2]).view(1, -1, 1)).any(-1).all()


def test_knn_canonical_tie_break():
    x = torch.zeros(1, 3, 8)
    x[0, 0] = torch.tensor([0., 1., 1., 2., 2., 3., 3., 3.])     # duplicated points -> exact ties
    idx = O.knn_canonical(x, 4)
    assert idx[0, 1].tolist() == [1, 2, 0, 3]                    # self tie -> lower index first
    assert idx[0, 2].tolist() == [1, 2, 0, 3]
    assert idx[0, 0].tolist() == [0, 1, 2, 3]


def test_graph_feature(golden):
    g = golden("edge_feat")
    x = t(g["x"]).requires_grad_(True)
    out = O.graph_feature(x, k=6)
    assert torch.equal(out, t(g["out"]))
    (gx,) = torch.autograd.grad(out, x, t(g["gout"]))
    torch.testing.assert_close(gx, t(g["gx"]), rtol=1e-5, atol=1e-5)
    xc = t(g["xc"]).requires_grad_(True)
    outc = O.graph_feature(xc, k=5, idx=t(g["idxc"], torch.int64), cross=True)
    torch.testing.assert_close(outc, t(g["outc"]), rtol=0, atol=0)
    (gxc,) = torch.autograd.grad(outc, xc, t(g["goutc"]))
    torch.testing.assert_close(gxc, t(g["gxc"]), rtol=1e-5, atol=1e-5)
    fixed = O.graph_feature(t(g["x"]), k=6, x_coord=t(g["coord"]))
    assert torch.equal(fixed, t(g["out_fixed"]))


@pytest.mark.parametrize("tag", ["s1e-3", "s1e-2", "s0.1", "s0.5", "s0.9", "mixed"])
def test_hyp_lca_fp64(golden, tag):
    g = golden("hyp_lca")
    a = t(g[tag + "_a"]).double().requires_grad_(True)
    b = t(g[tag + "_b"]).double().requires_grad_(True)
    dist = O.hyp_lca(a, b, return_coord=False)
    torch.testing.assert_close(dist, t(g[tag + "_dist"]), rtol=1e-9, atol=0)
    ga, gb = torch.autograd.grad(dist.sum(), (a, b))
    torch.testing.assert_close(ga, t(g[tag + "_ga"]), rtol=1e-7, atol=1e-9)
    torch.testing.assert_close(gb, t(g[tag + "_gb"]), rtol=1e-7, atol=1e-9)
    coord = O.hyp_lca(a, b, return_coord=True)
    torch.testing.assert_close(coord, t(g[tag + "_coord"]), rtol=1e-9, atol=1e-14)
    gca, gcb = torch.autograd.grad((coord * t(g[tag + "_gc"])).sum(), (a, b))
    torch.testing.assert_close(gca, t(g[tag + "_gca"]), rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(gcb, t(g[tag + "_gcb"]), rtol=1e-6, atol=1e-9)


def test_expmap_project(golden):
    g = golden("expmap")
    u = t(g["u"])
    assert torch.equal(O.expmap0(u), t(g["y"]))
    torch.testing.assert_close(O.expmap0(u.double()), t(g["y64"]), rtol=1e-14, atol=0)
    assert torch.equal(O.project(t(g["y"]) * 1.001), t(g["proj"]))


@pytest.mark.parametrize("frac", [0.0, 1.2])
def test_sampler_rng_parity(golden, frac):
    g = golden("compute_hyp")
    labels = t(g["labels"], torch.int64)
    torch.manual_seed(1234)
    a, p, n = O.sample_triplets(labels, t_per_anchor=7, fraction=frac)
    assert torch.equal(a, t(g[f"f{frac}_a"], torch.int64))
    assert torch.equal(p, t(g[f"f{frac}_p"], torch.int64))
    assert torch.equal(n, t(g[f"f{frac}_n"], torch.int64))
    assert (labels[a] == labels[p]).all() and (labels[a] != labels[n]).all() and (a != p).all()


@pytest.mark.parametrize("tag", ["s1e-3", "s0.1", "s0.5"])
@pytest.mark.parametrize("prec", ["64", "32"])
def test_compute_hyp(golden, tag, prec):
    g = golden("compute_hyp")
    dt = torch.float64 if prec == "64" else torch.float32
    x = t(g["x"]).to(dt).requires_grad_(True)
    scale = torch.tensor([float(g[tag + "_scale"])], dtype=dt, requires_grad=True)
    a, p, n = (t(g["f0.0_" + s], torch.int64) for s in "apn")
    a, p, n = O.filter_triplets(x.detach(), a, p, n, margin=0.0, kind="easy")
    assert a.numel() == int(g[f"{tag}_kept{prec}"])
    assert torch.equal(a, t(g[f"{tag}_kept_a{prec}"], torch.int64))
    loss = O.compute_hyp(x, a, p, n, scale, float(g[tag + "_temp"]))
    gx, gs = torch.autograd.grad(loss, (x, scale))
    if prec == "64":
        torch.testing.assert_close(loss, t(g[f"{tag}_loss64"]), rtol=1e-9, atol=0)   # s=1e-3 is ill-conditioned even in fp64
        torch.testing.assert_close(gx, t(g[f"{tag}_gx64"]), rtol=1e-6, atol=1e-10)
        torch.testing.assert_close(gs, t(g[f"{tag}_gscale64"]), rtol=1e-6, atol=1e-10)
    else:   # same ops, same torch build -> equal up to op-fusion differences
        torch.testing.assert_close(loss, t(g[f"{tag}_loss32"]), rtol=1e-6, atol=0)


@pytest.mark.parametrize("key", ["96", "200", "clu"])
@pytest.mark.parametrize("method,z", [("complete", "Zc"), ("single", "Zs")])
def test_decode_linkage(golden, key, method, z):
    g = golden("decode")
    Z = O.decode_linkage(t(g["x" + key]), torch.tensor([1e-3]), method=method)
    ref = g[z + key]
    assert np.array_equal(Z[:, [0, 1, 3]], ref[:, [0, 1, 3]])
    np.testing.assert_allclose(Z[:, 2], ref[:, 2], rtol=1e-15, atol=1e-16)


@pytest.mark.parametrize("method", ["single", "complete"])
@pytest.mark.parametrize("D", [32, 5])
def test_linkage_restatement_matches_scipy(method, D):
    from scipy.cluster.hierarchy import linkage
    from scipy.spatial.distance import pdist, squareform
    rng = np.random.default_rng(3)
    X = rng.standard_normal((60, D)).astype(np.float32)
    X[7] = X[3]                                        # exact duplicate -> distance ties
    X[20] = X[3]
    dm = O.pdist_cosine_restated(X)
    assert np.array_equal(dm, squareform(pdist(X.astype(np.float64), "cosine")))
    Z = O.linkage_restated(dm, method)
    assert np.array_equal(Z, linkage(X.astype(np.float64), method=method, metric="cosine"))


def test_fcluster_maxclust_restatement_matches_scipy():
    """The cut rule and cluster numbering csrc/cut.cu implements, restated in numpy, against scipy.fcluster itself:
    single and complete linkage, distinct and tied heights (duplicated points), k from 1 past N."""
    from scipy.cluster.hierarchy import fcluster, linkage
    rng = np.random.default_rng(5)
    for n, dup in ((3, False), (4, False), (7, True), (40, False), (40, True), (300, False), (300, True)):
        x = rng.standard_normal((n, 6))
        if dup:
            x[n // 2:n // 2 + n // 4] = x[:n // 4]                      # exact duplicates: tied heights
        for method in ("single", "complete"):
            Z = linkage(x, method=method, metric="euclidean")
            for k in list(range(1, min(n, 12))) + [n - 2, n - 1, n, n + 3]:
                if k < 1:
                    continue
                want = fcluster(Z, k, criterion="maxclust")
                got = O.fcluster_maxclust_restated(Z, k)
                assert np.array_equal(got, want), (n, dup, method, k)


def test_get_optimal_k_restatement_matches_reference_golden():
    """oracle.get_optimal_k_restated == the reference's get_optimal_k(y, Z, 'iou') run in the build container
    (tests/golden/optimal_k.npz, oracle/make_golden_cut.py): best partition, best k, score."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "optimal_k.npz"))
    for ci in range(int(g["n_cases"])):
        pred, k, score = O.get_optimal_k_restated(g[f"y{ci}"], g[f"Z{ci}"])
        assert k == int(g[f"k{ci}"]), ci
        assert np.array_equal(pred, g[f"pred{ci}"]), ci
        assert score == float(g[f"score{ci}"]), ci
        pred, k, score = O.get_optimal_k_restated(g[f"y{ci}"], g[f"Z{ci}"], index="ri")
        assert k == int(g[f"ri_k{ci}"]) and score == float(g[f"ri_score{ci}"]), ci
        if k > 0:
            assert np.array_equal(pred, g[f"ri_pred{ci}"]), ci


def test_fcluster_and_optimal_k_restatements_randomised():
    """Randomised sweep (seeded, no hypothesis shrinking needed): small dendrograms with many tied heights (integer
    lattices), every k from 1 past N, single and complete linkage: restated cut == scipy.fcluster, and the IoU / ARI
    model selection built on it == a direct evaluation with scipy + sklearn, the way the reference's get_optimal_k does."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from sklearn.metrics import adjusted_rand_score, jaccard_score
    rng = np.random.default_rng(123)
    for trial in range(40):
        n = int(rng.integers(3, 40))
        pts = rng.integers(0, 4, size=(n, 2)).astype(np.float64) if trial % 2 else rng.standard_normal((n, 3))
        Z = linkage(pts, method="single" if trial % 3 else "complete")
        for k in range(1, n + 3):
            assert np.array_equal(O.fcluster_maxclust_restated(Z, k), fcluster(Z, k, criterion="maxclust")), (trial, k)
        y = rng.integers(0, 4, n) * 2 + 1
        uniq = np.unique(y)
        yt = np.searchsorted(uniq, y)
        T = len(uniq)
        best_iou, best_ri = (0, 0.0), (0, 0.0)
        for k in range(1, T + 5):
            yp = fcluster(Z, k, criterion="maxclust") - 1
            P = len(np.unique(yp))
            m = np.zeros((T, P), dtype=np.float32)
            for i in range(T):
                for j in range(P):
                    m[i, j] = jaccard_score(yt == i, yp == j, zero_division=0)
            ind = m.argmax(1)
            remap = np.zeros_like(yp)
            for i in range(T):
                remap[yp == ind[i]] = i + 1
            a = np.eye(T + 1)[yt + 1]
            b = np.eye(T + 1)[remap]
            s_iou = np.logical_and(a, b).sum() / np.logical_or(a, b).sum()
            if s_iou > best_iou[1]:
                best_iou = (k, s_iou)
            s_ri = adjusted_rand_score(y, yp)
            if s_ri > best_ri[1]:
                best_ri = (k, s_ri)
        _, k1, s1 = O.get_optimal_k_restated(y, Z)
        _, k2, s2 = O.get_optimal_k_restated(y, Z, index="ri")
        assert (k1, s1) == best_iou, (trial, (k1, s1), best_iou)
        assert (k2, s2) == best_ri, (trial, (k2, s2), best_ri)


# ------------------------------------------------------------------------------------------------
# row f-1: the oracle's VN-layer restatement against the reference's own layers (tests/golden/edgeconv.npz)
# ------------------------------------------------------------------------------------------------
def _golden_convs(g, tag, dtype, with_buffers=True):
    convs, j = [], 0
    while f"{tag}_c{j}_wf" in g.files:
        c = {"wf": torch.tensor(g[f"{tag}_c{j}_wf"], dtype=dtype, requires_grad=True),
             "wd": torch.tensor(g[f"{tag}_c{j}_wd"], dtype=dtype, requires_grad=True),
             "gamma": torch.tensor(g[f"{tag}_c{j}_gamma"], dtype=dtype, requires_grad=True),
             "beta": torch.tensor(g[f"{tag}_c{j}_beta"], dtype=dtype, requires_grad=True)}
        if with_buffers:
            c["running_mean"] = torch.tensor(g[f"{tag}_c{j}_rm"], dtype=dtype)
            c["running_var"] = torch.tensor(g[f"{tag}_c{j}_rv"], dtype=dtype)
        convs.append(c)
        j += 1
    return convs


def _nrm_err(got, want):
    return ((got.double() - want.double()).norm() / want.double().norm().clamp_min(1e-300)).item()


@pytest.mark.parametrize("tag", ["l1", "l2", "l3"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_edgeconv_layer_restatement_vs_reference_layers(golden, tag, mode):
    """The oracle evaluated in fp64 equals the reference's layers evaluated in fp64 (1e-9); evaluated in fp32 it is as
    close to that as the reference's own fp32 evaluation is (which, for the C=1 layer, is only 1e-2 on some gradients)."""
    g = golden("edgeconv")
    idx = torch.tensor(g[f"{tag}_idx"]).long()
    names = ["gx"] + [f"c{j}_{n}" for j in range(2 if tag != "l3" else 1) for n in ("gwf", "gwd", "ggamma", "gbeta")]
    for dt in (torch.float64, torch.float32):
        convs = _golden_convs(g, tag, dt)
        x = torch.tensor(g[f"{tag}_x"]).to(dt).requires_grad_(True)
        y = O.edgeconv_layer(x, idx, convs, training=(mode == "train"))
        params = [c[n] for c in convs for n in ("wf", "wd", "gamma", "beta")]
        grads = torch.autograd.grad((y * torch.tensor(g[f"{tag}_gout"]).to(dt)).sum(), [x] + params)
        got = dict(zip(["y"] + names, [y.detach()] + list(grads)))
        for name, val in got.items():
            want64 = torch.tensor(g[f"{tag}_{mode}_{name}64"])
            if dt == torch.float64:
                assert _nrm_err(val, want64) < 1e-9, name
            else:
                ref32_err = _nrm_err(torch.tensor(g[f"{tag}_{mode}_{name}"]), want64)
                assert _nrm_err(val, want64) < max(1e-4, 3 * ref32_err), (name, ref32_err)
        if mode == "train" and dt == torch.float32:
            for j, c in enumerate(convs):
                assert torch.allclose(c["running_mean"], torch.tensor(g[f"{tag}_c{j}_rm_after"]), rtol=1e-5, atol=1e-6)
                assert torch.allclose(c["running_var"], torch.tensor(g[f"{tag}_c{j}_rv_after"]), rtol=1e-5, atol=1e-6)
