"""GPU drop-in tests: the native path driven THROUGH THE REFERENCE'S NAMES after ``hpcs_b200.patch.install()``.

``/root/reference`` does not exist on the GPU box, so the module tree here is the stand-in under
``tests/fake_reference`` (same module / class / attribute names and ``from ... import`` aliasing as the reference;
every hot-path function of it raises ``ReferencePathReached``).  ``tests/test_patch_host.py`` checks the identical
binding against the real reference on CPU.  Cases: a ShapeNet-shaped training step at BASELINE configs[1], the test
step (decode + model selection), a PartNet-shaped configs[2] step with the hierarchical loss subclass, the VN-PointNet
cross features, the --triplet-sim branch, and row f-4 (device input pipeline, replay-safe sampler state)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import hpcs_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-4
FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_reference")


@pytest.fixture(scope="module")
def tree():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    for m in [m for m in sys.modules if m == "hpcs" or m.startswith("hpcs.") or m == "train"]:
        del sys.modules[m]
    sys.path[:] = [p for p in sys.path if p != "/root/reference"]
    sys.path.insert(0, FAKE)
    import hpcs
    assert "fake_reference" in hpcs.__file__
    import hpcs.models, hpcs.nn.dgcnn, hpcs.nn.pointnet, hpcs.nn.hyperbolic   # noqa: F401,E401  (import BEFORE install, like train.py)
    import hpcs_b200.patch as patch
    patch.uninstall()
    done = patch.install(strict=True)
    assert patch.verify() == [] and len(done) >= 25
    yield hpcs
    patch.uninstall()
    sys.path.remove(FAKE)
    for m in [m for m in sys.modules if m == "hpcs" or m.startswith("hpcs.")]:
        del sys.modules[m]


def clouds(gen, B, N):
    pts = torch.randn(B, N, 3, generator=gen)
    pts = pts - pts.mean(1, keepdim=True)
    return pts / pts.norm(dim=-1).amax(1).view(B, 1, 1)            # [B,N,3], what the data loader yields


def shapenet_labels(gen, B, N, parts=(2, 6), classes=50):
    out = []
    for _ in range(B):
        n_parts = int(torch.randint(parts[0], parts[1] + 1, (1,), generator=gen))
        first = int(torch.randint(0, classes - n_parts + 1, (1,), generator=gen))
        probs = torch._sample_dirichlet(torch.ones(n_parts), generator=gen)
        out.append(torch.multinomial(probs, N, replacement=True, generator=gen) + first)
    return torch.stack(out)


def partnet_labels(gen, B, N, classes=39):
    out = []
    for _ in range(B):
        n_parts = int(torch.randint(3, 13, (1,), generator=gen))
        ids = torch.randperm(classes, generator=gen)[:n_parts]
        probs = torch._sample_dirichlet(torch.full((n_parts,), 0.3), generator=gen).clamp_min(1e-6)
        out.append(ids[torch.multinomial(probs / probs.sum(), N, replacement=True, generator=gen)])
    return torch.stack(out)


def rel_err(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-300)).item()


def launches():
    from hpcs_b200 import _lib
    return _lib.launch_count()


def oracle_step(model, pts, targets, mode, t_per_anchor, fraction, seed, x_poincare):
    """What the reference computes for the hyperbolic term, on CPU in fp64, with the same RNG consumption as the
    patched step: rotation draws first (shapenet_hyp_hc.py:64-67), then the sampler (loss_and_miner_utils.py)."""
    torch.manual_seed(seed)
    B = pts.shape[0]
    draws = torch.randn(B, 4) if mode == "so3" else torch.rand(B)
    R = O.quaternion_rotations(draws) if mode == "so3" else O.z_rotations(draws)
    labels = targets.reshape(-1)
    a, p, n = O.sample_triplets(labels, t_per_anchor, fraction)
    xd = x_poincare.detach().cpu().double()
    fa, fp_, fn_ = O.filter_triplets(xd, a, p, n, margin=0.0, kind="easy")
    loss = O.compute_hyp(xd, fa, fp_, fn_, model.scale.detach().cpu().double(), model.temperature)
    return O.rotate_points(pts, R), loss


@pytest.mark.parametrize("B,full", [(4, False), (32, True)])
def test_shapenet_training_step_through_patched_names(tree, B, full):
    """BASELINE configs[1] (B=32, N=1024, k=20, 32-d embeddings, 50 triplets per anchor, fraction 0) and a 4-cloud
    version of it small enough for the dense fp64 oracle."""
    hpcs = tree
    from hpcs.models import ShapeNetHypHC
    from hpcs.nn.dgcnn import VN_DGCNN_partseg
    from hpcs.nn.hyperbolic import ExpMap
    N, k = 1024, 20
    gen = torch.Generator().manual_seed(100 + B)
    torch.manual_seed(1)
    model = ShapeNetHypHC(nn_feat=VN_DGCNN_partseg(3, 32, k, 0.0, "mean", 16), nn_emb=ExpMap(), euclidean_size=32,
                          hyp_size=32, num_class=50, t_per_anchor=50, fraction=0.0, temperature=0.05, miner=True).cuda()
    pts, targets = clouds(gen, B, N), shapenet_labels(gen, B, N)
    label = torch.randint(0, 16, (B, 1), generator=gen)
    seen = {}
    state0 = {k_: v.detach().clone() for k_, v in model.nn_feat.state_dict().items()}
    def keep(mod, inp, out):
        seen["x_poincare"], seen["x_euclidean"] = out, inp[0]
    model.nn_emb.register_forward_hook(keep)
    before = launches()
    torch.manual_seed(7)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False              # the dense tail's Conv1d layers in fp32, like the host evaluation below
    try:
        losses, metrics = model.forward((pts, label, targets), testing=False)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    total = losses["loss_metric"] + losses["loss_hyp"]
    total.backward()
    torch.cuda.synchronize()
    assert launches() - before >= 3 + 8 + 2 + 5          # >= 3 kNN, 8 fused-layer forward passes, loss fwd+bwd, 5 fused backward kernels
    assert torch.isfinite(total).item()
    for name, prm in model.named_parameters():
        fused = any(f"nn_feat.conv{i}." in name for i in range(1, 6))          # parameters of the five graph-layer convs
        assert prm.grad is not None or not (fused or name == "scale"), name   # (stand-in tail modules own unused ones)
        assert prm.grad is None or torch.isfinite(prm.grad).all(), name
        assert not fused or prm.grad.abs().max().item() > 0, name
    assert model.scale.grad.abs().item() > 0
    xp = seen["x_poincare"].reshape(-1, 32)
    assert tuple(xp.shape) == (B * N, 32) and xp.norm(dim=1).max().item() <= 1.0 + 1e-6
    if not full:
        rotated, want = oracle_step(model, pts, targets, "so3", 50, 0.0, 7, xp)
        assert abs(losses["loss_hyp"].item() / model.trade_off - want.item()) <= REL * abs(want.item())
        # the rotation + transpose the backbone saw (row f-4), its three fused graph layers (row f-1; kNN D=3, 63, 63) and
        # the dense tail: the same forward on the host with the oracle's layers standing in for the fused ones
        import copy
        import hpcs_b200.edgeconv as ec
        net = model.nn_feat
        cpu_net = copy.deepcopy(net).cpu()
        cpu_net.load_state_dict({k_: v.cpu() for k_, v in state0.items()})        # parameters / buffers as before the step

        def edgeconv_oracle(x, k_, conv_a, conv_b=None, idx=None):
            Bc, C, _, Np = x.shape
            convs = []
            for c in (conv_a, conv_b):
                if c is not None:
                    bn = c.batchnorm.bn
                    convs.append({"wf": c.map_to_feat.weight, "wd": c.map_to_dir.weight, "gamma": bn.weight, "beta": bn.bias})
            return O.edgeconv_layer(x, O.knn_canonical(x.reshape(Bc, 3 * C, Np).contiguous(), k_), convs, training=True)
        real = ec.edgeconv
        ec.edgeconv = edgeconv_oracle
        try:
            with torch.no_grad():
                want_feat = ec.vn_dgcnn_partseg_forward(cpu_net, rotated, O.to_categorical(label, 16))
        finally:
            ec.edgeconv = real
        got_feat = seen["x_euclidean"].detach().cpu()
        # near-tied neighbours may swap between the fp32 GPU and host evaluations of layers 2 and 3; everything else agrees
        # (one swapped neighbour moves that point's feature by O(1/k) and, through BatchNorm and the global max, nudges the rest)
        rel = (got_feat - want_feat).norm(dim=-1) / want_feat.norm(dim=-1).clamp_min(1e-6)
        assert rel.median().item() < 5e-4 and (rel <= 2e-3).float().mean().item() > 0.9, (rel.median().item(), (rel <= 2e-3).float().mean().item())


def test_test_step_decode_and_model_selection_through_patched_names(tree):
    """forward(testing=True) (base_hyp_hc.py:120-140; the binding replaces its per-cloud ``_decode_linkage`` loop,
    :133-137, by one batched decode) and ``get_optimal_k`` (:197-199) through the names ``base_hyp_hc`` imported --
    same return tuple, Z bit-equal to scipy per cloud, same best k / score; ``_decode_linkage`` called directly agrees."""
    hpcs = tree
    from scipy.cluster.hierarchy import linkage
    from hpcs.models import ShapeNetHypHC
    from hpcs.nn.dgcnn import VN_DGCNN_partseg
    from hpcs.nn.hyperbolic import ExpMap
    B, N = 3, 256
    gen = torch.Generator().manual_seed(5)
    torch.manual_seed(2)
    model = ShapeNetHypHC(nn_feat=VN_DGCNN_partseg(3, 32, 10, 0.5, "mean", 16), nn_emb=ExpMap(), euclidean_size=32,
                          hyp_size=32, num_class=50, t_per_anchor=10, fraction=0.0, test_rotation="z").cuda()
    pts, targets = clouds(gen, B, N), shapenet_labels(gen, B, N)
    label = torch.randint(0, 16, (B, 1), generator=gen)
    with torch.no_grad():
        losses, _, x_e, x_p, Zs, points, labels = model.forward((pts, label, targets), testing=True)
    assert len(Zs) == B and tuple(points.shape) == (B, 3, N)
    from hpcs_b200.hyperbolic import normalize_project
    for i in range(B):
        assert isinstance(Zs[i], np.ndarray) and Zs[i].dtype == np.float64 and Zs[i].shape == (N - 1, 4)
        leaves = normalize_project(x_p[i], model.scale).cpu().numpy()
        assert np.array_equal(Zs[i], linkage(leaves, method="complete", metric="cosine"))
        assert np.array_equal(model._decode_linkage(x_p[i]), Zs[i])
    assert set(losses) == {"loss_metric", "loss_hyp"}
    scores = model.test_scores(labels, Zs)
    for i, (pred, kbest, score) in enumerate(scores):
        wpred, wk, wscore = O.get_optimal_k_restated(labels[i].cpu().numpy(), Zs[i])
        assert kbest == wk and score == pytest.approx(wscore, abs=1e-12)
        assert np.array_equal(pred, wpred)


@pytest.mark.parametrize("B,full", [(2, False), (8, True)])
def test_partnet_hierarchical_step_through_patched_names(tree, B, full):
    """BASELINE configs[2], one rank's share: PartNet-Chair-shaped (39 classes, 4-d embeddings, fraction 1.2,
    Dirichlet(0.3) part sizes, MLPExpMap).  ``--hierarchical`` is on by default (train.py:53), so the loss object is the
    SUBCLASS ``HierarchicalMetricHyperbolicLoss``; its hyperbolic term must be the fused kernel, not the dense matrix."""
    hpcs = tree
    from hpcs.models import PartNetHypHC
    from hpcs.nn.dgcnn import VN_DGCNN_partseg
    from hpcs.nn.hyperbolic import MLPExpMap
    import hpcs.loss.ultrametric_loss as ul
    N, k = 1024, 20
    gen = torch.Generator().manual_seed(200 + B)
    torch.manual_seed(3)
    hier = [[list(range(0, 13)), list(range(13, 26)), list(range(26, 39))], [[c] for c in range(39)]]
    model = PartNetHypHC(nn_feat=VN_DGCNN_partseg(3, 4, k, 0.5, "mean", 1), nn_emb=MLPExpMap(4, 4), euclidean_size=4,
                         hyp_size=4, num_class=39, t_per_anchor=50, fraction=1.2, temperature=0.1, miner=True,
                         hierarchical=True, hierarchy_list=hier).cuda()
    assert type(model.metric_hyp_loss) is ul.HierarchicalMetricHyperbolicLoss
    pts, targets = clouds(gen, B, N), partnet_labels(gen, B, N)
    seen = {}
    model.nn_emb.register_forward_hook(lambda mod, inp, out: seen.__setitem__("x_poincare", out))
    torch.manual_seed(9)
    losses, _ = model.forward((pts, targets), testing=False)
    (losses["loss_metric"] + losses["loss_hyp"]).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(losses["loss_hyp"]).item() and torch.isfinite(losses["loss_metric"]).item()
    assert model.scale.grad is not None and torch.isfinite(model.scale.grad).all()
    assert all(torch.isfinite(p.grad).all() for p in model.nn_emb.parameters())
    if not full:
        xp = seen["x_poincare"].reshape(-1, 4)
        _, want = oracle_step(model, pts, targets, "so3", 50, 1.2, 9, xp)
        assert abs(losses["loss_hyp"].item() / model.trade_off - want.item()) <= REL * abs(want.item())


def test_vn_pointnet_cross_features_through_patched_names(tree):
    hpcs = tree
    from hpcs.nn.pointnet import VN_POINTNET_partseg
    gen = torch.Generator().manual_seed(6)
    x = clouds(gen, 3, 200).transpose(1, 2).contiguous()
    net = VN_POINTNET_partseg(k=12)
    got = net(x.cuda()).cpu()
    want = O.graph_feature(x.unsqueeze(1), k=12, cross=True).mean(dim=-1)
    assert tuple(got.shape) == (3, 3, 3, 200)
    assert torch.allclose(got, want, atol=1e-6)


def test_triplet_sim_branch_through_patched_names(tree):
    """--triplet-sim (cosface=False): the reference's miner object and its TripletMarginLoss object, native bodies."""
    hpcs = tree
    import hpcs.loss.ultrametric_loss as ul
    gen = torch.Generator().manual_seed(8)
    n, D = 600, 16
    x = O.expmap0(torch.randn(n, D, generator=gen) * 0.5)
    labels = torch.randint(0, 5, (n,), generator=gen)
    scale = torch.nn.Parameter(torch.tensor([0.1]))
    mod = ul.MetricHyperbolicLoss(margin=0.05, t_per_anchor=8, fraction=1.2, scale=scale, temperature=0.05, num_class=5,
                                  embedding_size=D, cosface=False, miner=True).cuda()
    xg = x.cuda().requires_grad_(True)
    torch.manual_seed(12)
    out = mod.compute_loss(xg, xg, labels.cuda())
    torch.manual_seed(12)
    xd = x.double()
    a, p, ng = O.sample_triplets(labels, 8, 1.2)
    want_hyp = O.compute_hyp(xd, *O.filter_triplets(xd, a, p, ng, 0.0, "easy"), torch.tensor([0.1], dtype=torch.float64), 0.05)
    a2, p2, n2 = O.sample_triplets(labels, 8, 1.2)                  # the metric term mines again (ultrametric_loss.py:121)
    fa, fp_, fn_ = O.filter_triplets(xd, a2, p2, n2, 0.05, "semihard")
    sim = O.cosine_similarity_matrix(xd)
    viol = torch.relu(sim[fa, fn_] - sim[fa, fp_] + 0.05)
    want_metric = viol[viol > 0].mean()
    assert abs(out["loss_hyp"]["losses"].item() - want_hyp.item()) <= REL * abs(want_hyp.item())
    assert abs(out["loss_metric"]["losses"].item() - want_metric.item()) <= REL * abs(want_metric.item())
    (out["loss_hyp"]["losses"] + out["loss_metric"]["losses"]).backward()
    assert torch.isfinite(xg.grad).all()


def test_unpatched_tree_raises(tree):
    """The stand-in really is inert: without install() the same call reaches ReferencePathReached."""
    import hpcs
    import hpcs_b200.patch as patch
    from hpcs.nn.dgcnn import VN_DGCNN_partseg
    patch.uninstall()
    try:
        net = VN_DGCNN_partseg(3, 8, 4, 0.5, "mean", 16).cuda()
        with pytest.raises(hpcs.ReferencePathReached):
            net(torch.randn(2, 3, 32).cuda(), torch.zeros(2, 16, 1).cuda())
    finally:
        patch.install(strict=True)
    assert patch.verify() == []


# ------------------------------------------------------------------------------------------------
# row f-4: input pipeline on the device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["so3", "z", None, "matrix"])
def test_rotate_points_vs_oracle(mode):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import hpcs_b200 as hb
    gen = torch.Generator().manual_seed(3)
    B, N = 5, 1000
    pts = clouds(gen, B, N)
    if mode == "so3":
        params = torch.randn(B, 4, generator=gen)
        params[1, 0] = -abs(params[1, 0])                 # negative real part: the copysign branch
        R = O.quaternion_rotations(params)
    elif mode == "z":
        params = torch.rand(B, generator=gen)
        R = O.z_rotations(params)
    elif mode == "matrix":
        params = R = O.quaternion_rotations(torch.randn(B, 4, generator=gen))
    else:
        params = R = None
    got, rot = hb.rotate_points(pts.cuda(), mode, params, return_rotation=True)
    want = O.rotate_points(pts, R)
    assert tuple(got.shape) == (B, 3, N) and got.is_contiguous()
    assert torch.allclose(got.cpu(), want, atol=2e-6)
    if R is not None:
        assert torch.allclose(rot.cpu(), R, atol=2e-6)
        eye = torch.eye(3).expand(B, 3, 3)
        assert torch.allclose(rot.cpu() @ rot.cpu().transpose(1, 2), eye, atol=1e-5)
        assert torch.allclose(torch.det(rot.cpu()), torch.ones(B), atol=1e-5)
        assert torch.allclose(got.cpu().norm(dim=1), pts.norm(dim=2), atol=1e-5)      # rotations keep lengths
    else:
        assert torch.equal(got.cpu(), want)
    # host draws in the reference's order: same seed -> same rotation as pytorch3d's call would give
    if mode in ("so3", "z"):
        torch.manual_seed(44)
        auto = hb.rotate_points(pts.cuda(), mode)
        torch.manual_seed(44)
        draws = torch.randn(B, 4) if mode == "so3" else torch.rand(B)
        Rw = O.quaternion_rotations(draws) if mode == "so3" else O.z_rotations(draws)
        assert torch.allclose(auto.cpu(), O.rotate_points(pts, Rw), atol=2e-6)


def test_to_categorical_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import hpcs_b200 as hb
    y = torch.randint(0, 16, (32, 1))
    got = hb.to_categorical(y.cuda(), 16)
    assert got.is_cuda and torch.equal(got.cpu(), O.to_categorical(y, 16))
    assert torch.equal(hb.to_categorical(torch.zeros(4, 1, dtype=torch.long).cuda(), 1).cpu(), torch.ones(4, 1, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hb.to_categorical(y, 16)


def test_sampler_state_advances_under_graph_replay():
    """ROUND-1 verdict, weak #7: a by-value seed is frozen into a captured graph.  With the key in device memory every
    replay draws new triplets, and restoring the state reproduces them."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import hpcs_b200 as hb
    gen = torch.Generator().manual_seed(0)
    labels = torch.randint(0, 7, (4096,), generator=gen)
    order, seg, T0 = hb.triplet_plan(labels, 5, 0.0)
    plan = (order.cuda(), seg.cuda(), T0)
    state = hb.sampler_state(seed=123)
    out = hb.sample_triplets_device(None, plan=plan, state=state)           # warm-up outside the graph
    torch.cuda.synchronize()
    assert state.cpu().tolist() == [123, 1, 0]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = hb.sample_triplets_device(None, plan=plan, state=state)
    draws = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        draws.append(tuple(t.clone() for t in out))
    assert state.cpu().tolist() == [123, 4, 0]
    lab = labels.cuda()
    for a, p, n in draws:
        assert torch.equal(a, draws[0][0])                                   # anchors are deterministic
        assert (lab[a.long()] == lab[p.long()]).all() and (lab[a.long()] != lab[n.long()]).all() and (a != p).all()
    assert not torch.equal(draws[0][1], draws[1][1]) and not torch.equal(draws[1][1], draws[2][1])
    state.copy_(torch.tensor([123, 2, 0], dtype=torch.int64))              # rewind: the second replay again
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(out, draws[1]))


@pytest.mark.parametrize("n,D,classes", [(1000, 32, 50), (8192, 4, 39), (65, 16, 1), (3, 7, 130)])
def test_cosface_logits_kernel_vs_oracle(n, D, classes):
    """``hpcs_cosface_logits_f32`` against the restated ``get_logits`` (ultrametric_loss.py:95-112): ragged row counts,
    a zero embedding row (F.normalize's eps), PartNet's 4-d embeddings."""
    from hpcs_b200.loss import cosface_logits
    gen = torch.Generator().manual_seed(n + classes)
    emb = torch.randn(n, D, generator=gen) * 0.3
    emb[0] = 0.0
    W = torch.randn(D, classes, generator=gen)
    y = torch.randint(0, classes, (n,), generator=gen)
    got = cosface_logits(emb.cuda(), W.cuda(), y.cuda(), 0.35, 64.0).cpu()
    want = O.cosface_logits(emb.double(), W.double(), y, 0.35, 64.0)
    assert got.shape == (n, classes)
    assert (got.double() - want).abs().max().item() < 1e-4 * 64.0


def test_get_logits_through_patched_names_is_memoised(tree):
    """``MetricHyperbolicLoss.get_logits`` (bound method of the reference-side class): equal to the reference formula, the
    second call of a step with the same tensors returns the first result, an in-place weight update invalidates it."""
    hpcs = tree
    from hpcs.loss.ultrametric_loss import MetricHyperbolicLoss
    torch.manual_seed(3)
    loss = MetricHyperbolicLoss(num_class=50, embedding_size=32, cosface=True, miner=True).cuda()
    x = (torch.randn(4096, 32) * 0.2).cuda()
    y = torch.randint(0, 50, (4096,)).cuda()
    a = loss.get_logits(x, y)
    lc = loss.loss_cosface
    want = O.cosface_logits(x.cpu().double(), lc.W.detach().cpu().double(), y.cpu(), lc.margin, lc.scale)
    assert (a.cpu().double() - want).abs().max().item() < 1e-4 * lc.scale
    assert loss.get_logits(x, y) is a
    with torch.no_grad():
        lc.W.mul_(-1.0)
    b = loss.get_logits(x, y)
    assert b is not a and not torch.equal(a, b)
    x2 = x.clone()
    assert loss.get_logits(x2, y) is not b


def test_device_sampler_binding(tree):
    """``install(sampler='device')``: the bound ``compute_hyp`` draws its triplets on the GPU (row f-3) -- same anchors in the
    same order as the reference sampler, fresh draws per call, loss close to the host-sampled one (both are Monte-Carlo
    estimates over 50 triplets per anchor of the same quantity); the default binding is put back afterwards."""
    hpcs = tree
    import hpcs_b200.patch as patch
    from hpcs_b200 import loss as L
    from hpcs.loss.ultrametric_loss import MetricHyperbolicLoss
    torch.manual_seed(11)
    gen = torch.Generator().manual_seed(11)
    obj = MetricHyperbolicLoss(num_class=50, embedding_size=32, cosface=True, miner=True, t_per_anchor=50, fraction=0.0,
                               scale=torch.nn.Parameter(torch.tensor([1e-3]))).cuda()
    labels = shapenet_labels(gen, 4, 512).reshape(-1).cuda()
    u = torch.randn(4 * 512, 32, generator=gen)
    x = (torch.tanh(u.norm(dim=-1, keepdim=True)) * u / u.norm(dim=-1, keepdim=True)).cuda().requires_grad_(True)
    try:
        patch.install(strict=True, sampler="device")
        assert L.BOUND_SAMPLER == "device"
        before = launches()
        l1 = obj.compute_hyp(x, labels)
        l2 = obj.compute_hyp(x, labels)
        assert launches() - before >= 2 * 2
        g = torch.autograd.grad(l1, x)[0]
        assert torch.isfinite(g).all() and g.abs().max().item() > 0
        assert l1.item() != l2.item()                                   # the key advances on the device: fresh triplets
        a_dev = L._bound_sample(obj, labels, 50, 0.0)[0].long().cpu()
        patch.install(strict=True, sampler="reference")
        a_ref = L._bound_sample(obj, labels, 50, 0.0)[0].cpu()
        assert torch.equal(a_dev, a_ref)                                # anchors: identical order
        l_ref = obj.compute_hyp(x, labels)
        assert abs(l1.item() - l_ref.item()) < 0.05 * abs(l_ref.item())
    finally:
        patch.install(strict=True, sampler="reference")
    assert L.BOUND_SAMPLER == "reference"
