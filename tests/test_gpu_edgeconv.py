"""GPU parity of the fused EdgeConv layer (csrc/edgeconv.cu) through the C ABI:

* against ``tests/golden/edgeconv.npz`` -- the reference's own ``get_graph_feature`` + ``VNLinearLeakyReLU`` (+ second one)
  + ``mean_pool`` evaluated in fp64 by the unmodified reference (oracle/make_golden_layers.py), training and eval mode:
  output, gradient wrt the input and wrt every parameter, BatchNorm running buffers;
* against the fp64 oracle restatement at larger shapes (k = 10 / 20 / 40, ragged tiles);
* at BASELINE's full layer shape (B=32, N=1024, k=20, C=21) against the unfused native path (edge-feature kernel + the same
  VN arithmetic in plain PyTorch on the device).
Tolerance: 1e-4 norm-wise (north_star's fp32 bar).  Exception, stated where it applies: gradients of the C=1 layer, where
the reference's own fp32 evaluation is up to 2e-2 away from its fp64 one (a vector channel with |p| ~ 0 on some edge has a
1/|p| gradient); there the bar is 3x the reference's own fp32 error recorded in the fixture."""
import numpy as np
import pytest
import torch

from oracle import hpcs_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-4


@pytest.fixture(scope="module")
def hb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import hpcs_b200
    return hpcs_b200


def nrm_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-300)).item()


class VNConv(torch.nn.Module):
    """Attribute layout of the reference's VNLinearLeakyReLU (vn_layers.py:48-60): what hpcs_b200.edgeconv reads."""

    def __init__(self, cin, cout=21):
        super().__init__()
        self.negative_slope = 0.2
        self.map_to_feat = torch.nn.Linear(cin, cout, bias=False)
        self.map_to_dir = torch.nn.Linear(cin, cout, bias=False)
        self.batchnorm = torch.nn.Module()
        self.batchnorm.bn = torch.nn.BatchNorm2d(cout)

    def as_oracle(self, dtype=torch.float64, grad=True):
        bn = self.batchnorm.bn
        d = {"wf": self.map_to_feat.weight, "wd": self.map_to_dir.weight, "gamma": bn.weight, "beta": bn.bias}
        d = {k: v.detach().cpu().to(dtype).requires_grad_(grad) for k, v in d.items()}
        d.update(running_mean=bn.running_mean.detach().cpu().to(dtype).clone(), running_var=bn.running_var.detach().cpu().to(dtype).clone(),
                 eps=bn.eps, momentum=bn.momentum)
        return d


def convs_from_golden(g, tag):
    convs, j = [], 0
    while f"{tag}_c{j}_wf" in g.files:
        wf = torch.tensor(g[f"{tag}_c{j}_wf"])
        c = VNConv(wf.shape[1])
        with torch.no_grad():
            c.map_to_feat.weight.copy_(wf)
            c.map_to_dir.weight.copy_(torch.tensor(g[f"{tag}_c{j}_wd"]))
            c.batchnorm.bn.weight.copy_(torch.tensor(g[f"{tag}_c{j}_gamma"]))
            c.batchnorm.bn.bias.copy_(torch.tensor(g[f"{tag}_c{j}_beta"]))
            c.batchnorm.bn.running_mean.copy_(torch.tensor(g[f"{tag}_c{j}_rm"]))
            c.batchnorm.bn.running_var.copy_(torch.tensor(g[f"{tag}_c{j}_rv"]))
        convs.append(c.cuda())
        j += 1
    return convs


def param_list(convs):
    return [p for c in convs for p in (c.map_to_feat.weight, c.map_to_dir.weight, c.batchnorm.bn.weight, c.batchnorm.bn.bias)]


@pytest.mark.parametrize("tag", ["l1", "l2", "l3"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_fused_layer_vs_reference_layers_golden(hb, golden, tag, mode):
    from hpcs_b200.edgeconv import edgeconv
    g = golden("edgeconv")
    convs = convs_from_golden(g, tag)
    for c in convs:
        c.train(mode == "train")
    x = torch.tensor(g[f"{tag}_x"]).cuda().requires_grad_(True)
    idx = torch.tensor(g[f"{tag}_idx"]).long().cuda()
    k = idx.shape[2]
    y = edgeconv(x, k, convs[0], convs[1] if len(convs) == 2 else None, idx=idx)
    assert tuple(y.shape) == (x.shape[0], 21, 3, x.shape[3])
    grads = torch.autograd.grad((y * torch.tensor(g[f"{tag}_gout"]).cuda()).sum(), [x] + param_list(convs))
    names = ["gx"] + [f"c{j}_{n}" for j in range(len(convs)) for n in ("gwf", "gwd", "ggamma", "gbeta")]
    got = dict(zip(["y"] + names, [y] + list(grads)))
    for name, val in got.items():
        want64 = torch.tensor(g[f"{tag}_{mode}_{name}64"])
        ref32_err = nrm_err(torch.tensor(g[f"{tag}_{mode}_{name}"]), want64)        # the reference's own fp32 deviation
        bar = REL if tag != "l1" or name == "y" else max(REL, 3 * ref32_err)
        assert nrm_err(val, want64) < bar, (name, nrm_err(val, want64), ref32_err)
    if mode == "train":
        for j, c in enumerate(convs):
            bn = c.batchnorm.bn
            assert torch.allclose(bn.running_mean.cpu(), torch.tensor(g[f"{tag}_c{j}_rm_after"]), rtol=1e-5, atol=1e-6)
            assert torch.allclose(bn.running_var.cpu(), torch.tensor(g[f"{tag}_c{j}_rv_after"]), rtol=1e-5, atol=1e-6)
            assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize("B,C,N,k,two,train", [(3, 21, 200, 20, True, True), (2, 21, 333, 10, True, False), (2, 1, 256, 20, True, True),
                                                (2, 21, 150, 40, False, True), (1, 21, 77, 7, True, True), (4, 21, 128, 20, False, False),
                                                (2, 5, 130, 12, True, True), (1, 2, 64, 8, False, True)])
def test_fused_layer_vs_oracle_fp64(hb, B, C, N, k, two, train):
    layer_case(hb, B, C, N, k, two, train)


def layer_case(hb, B, C, N, k, two, train, conditioning_aware=False):
    """One fused layer against the fp64 oracle.  ``conditioning_aware`` (tools/stress_edgeconv.py): the layer is not
    continuous -- the direction test ``dot >= 0`` switches a term of the gradient on and off, and ``|p| -> 0`` has a
    ``1/|p|`` gradient -- so over random shapes a few per cent of the layers hold an (edge, channel) pair within fp32
    rounding of a discontinuity, and ANY fp32 evaluation is then 1e-4..1e-3 off fp64 in some gradient.  The bar of a gradient
    becomes max(1e-4, 3 x the largest change of the fp64 ORACLE's own gradient when its inputs and weights are perturbed by
    1e-6 relative, three draws): it stays 1e-4 wherever the reference function is smooth at fp32 resolution."""
    from hpcs_b200.edgeconv import edgeconv
    gen = torch.Generator().manual_seed(B * 1000 + N + k)
    torch.manual_seed(N + k)
    convs = [VNConv(2 * C)] + ([VNConv(21)] if two else [])
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
    grads = torch.autograd.grad((y * gout.cuda()).sum(), [xg] + param_list(convs))
    xd = x.double().requires_grad_(True)
    yw = O.edgeconv_layer(xd, idx.cpu(), oc, training=train)
    wparams = [c[n] for c in oc for n in ("wf", "wd", "gamma", "beta")]
    wgrads = torch.autograd.grad((yw * gout.double()).sum(), [xd] + wparams)
    assert nrm_err(y, yw) < REL
    bars = [REL] * len(wgrads)
    if C == 1:          # coordinates layer: a channel with |p| ~ 0 on some edge has a 1/|p| gradient; the bar is what the same
        o32 = [c.as_oracle(torch.float32) for c in convs]      # formulas give when evaluated in fp32 on the host (x5)
        x32 = x.clone().requires_grad_(True)
        y32 = O.edgeconv_layer(x32, idx.cpu(), o32, training=train)
        g32 = torch.autograd.grad((y32 * gout).sum(), [x32] + [c[n] for c in o32 for n in ("wf", "wd", "gamma", "beta")])
        bars = [max(REL, 5 * nrm_err(a, b)) for a, b in zip(g32, wgrads)]
    if conditioning_aware:
        pg = torch.Generator().manual_seed(99)
        worst = [0.0] * len(wgrads)
        for _ in range(3):
            jit = lambda t: (t.detach() * (1 + 1e-6 * torch.randn(t.shape, generator=pg, dtype=t.dtype))).requires_grad_(True)   # noqa: E731
            op = [dict(c, wf=jit(c["wf"]), wd=jit(c["wd"]), gamma=jit(c["gamma"]), beta=jit(c["beta"]),
                       running_mean=c["running_mean"].clone(), running_var=c["running_var"].clone()) for c in [cv.as_oracle() for cv in convs]]
            xp = jit(x.double())
            yp = O.edgeconv_layer(xp, idx.cpu(), op, training=train)
            gp = torch.autograd.grad((yp * gout.double()).sum(), [xp] + [c[n] for c in op for n in ("wf", "wd", "gamma", "beta")])
            worst = [max(w, nrm_err(a, b)) for w, a, b in zip(worst, gp, wgrads)]
        bars = [max(b, 3 * w) for b, w in zip(bars, worst)]
    for i, (gg, wg) in enumerate(zip(grads, wgrads)):
        assert nrm_err(gg, wg) < bars[i], (i, nrm_err(gg, wg), bars[i])
    if train:
        for c, o in zip(convs, oc):
            assert torch.allclose(c.batchnorm.bn.running_mean.cpu().double(), o["running_mean"], rtol=1e-5, atol=1e-6)
            assert torch.allclose(c.batchnorm.bn.running_var.cpu().double(), o["running_var"], rtol=1e-5, atol=1e-6)


def test_fused_layer_full_shape_vs_unfused_native_path(hb):
    """B=32, N=1024, k=20, C=21, two convs, training mode: the fused layer against the unfused composition on the GPU
    (hpcs_b200.get_graph_feature -> the same VN arithmetic in PyTorch fp32 -> mean), forward and backward."""
    from hpcs_b200.edgeconv import edgeconv
    B, C, N, k = 32, 21, 1024, 20
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    convs = [VNConv(2 * C).cuda().train(), VNConv(21).cuda().train()]
    x = torch.randn(B, C, 3, N, generator=gen).cuda()
    gout = torch.randn(B, 21, 3, N, generator=gen).cuda()
    idx = hb.knn(x.view(B, 3 * C, N), k)
    xa = x.clone().requires_grad_(True)
    y = edgeconv(xa, k, convs[0], convs[1], idx=idx)
    ga = torch.autograd.grad((y * gout).sum(), [xa] + param_list(convs))
    assert torch.isfinite(y).all()
    xb = x.clone().requires_grad_(True)
    e = hb.get_graph_feature(xb, k, idx=idx)
    ob = []
    for c in convs:
        o = c.as_oracle(torch.float32, grad=False)
        o = {k_: (v.cuda().requires_grad_(True) if k_ in ("wf", "wd", "gamma", "beta") else v) for k_, v in o.items()}
        o["running_mean"] = o["running_var"] = None
        ob.append(o)
        e = O.vn_linear_leaky_relu(e, o["wf"], o["wd"], o["gamma"], o["beta"], None, None, True)
    yb = e.mean(dim=-1)
    gb = torch.autograd.grad((yb * gout).sum(), [xb] + [o[n] for o in ob for n in ("wf", "wd", "gamma", "beta")])
    assert nrm_err(y, yb) < REL
    for i, (p, q) in enumerate(zip(ga, gb)):
        assert nrm_err(p, q) < 2 * REL, (i, nrm_err(p, q))            # both sides are fp32 here


def test_fused_backbone_forward_matches_unfused(hb):
    """``vn_dgcnn_partseg_forward`` (what patch.install binds onto VN_DGCNN_partseg) on a stand-in module with the
    reference's attribute names: same output as running the three graph layers unfused."""
    from hpcs_b200.edgeconv import edgeconv, vn_dgcnn_partseg_forward

    class Tail(torch.nn.Module):
        def forward(self, x):
            return x

    class Backbone(torch.nn.Module):
        pooling, k = "mean", 12

        def __init__(self):
            super().__init__()
            self.conv1, self.conv2 = VNConv(2), VNConv(21)
            self.conv3, self.conv4, self.conv5 = VNConv(42), VNConv(21), VNConv(42)
            self.conv6 = Tail()
            self.std_feature = lambda f: (f, torch.eye(3, device=f.device).view(1, 3, 3, 1).expand(f.shape[0], 3, 3, f.shape[-1]))
            self.conv7 = torch.nn.Conv1d(16, 8, 1)
            self.conv8 = torch.nn.Conv1d(63 * 2 * 3 + 8 + 63 * 3, 16, 1)
            self.dp1 = self.dp2 = self.conv9 = self.conv10 = Tail()
            self.conv11 = torch.nn.Conv1d(16, 5, 1)

    torch.manual_seed(3)
    net = Backbone().cuda().eval()
    x = torch.randn(2, 3, 96).cuda() + 2.0
    l = torch.zeros(2, 16, 1).cuda()
    out = vn_dgcnn_partseg_forward(net, x, l)
    assert tuple(out.shape) == (2, 96, 5) and torch.isfinite(out).all()
    x1 = edgeconv(x.unsqueeze(1), 12, net.conv1, net.conv2)
    e = hb.get_graph_feature(x.unsqueeze(1), 12)
    for c in (net.conv1, net.conv2):
        o = c.as_oracle(torch.float32, grad=False)
        e = O.vn_linear_leaky_relu(e, o["wf"].cuda(), o["wd"].cuda(), o["gamma"].cuda(), o["beta"].cuda(), o["running_mean"].cuda(),
                                   o["running_var"].cuda(), False)
    assert nrm_err(x1, e.mean(dim=-1)) < REL
