"""CPU checks of the plain-PyTorch members of the host mirror (outside the CUDA hot path but part of the drop-in
surface): CosineSimilarity, CosFaceLoss, HierarchicalCosFaceLoss, the triplet-margin branch, get_triplets,
anneal_temperature -- against the reference's own classes where /root/reference is present, else against the
oracle's restatements."""
import os
import sys

import pytest
import torch

import hpcs_b200 as hb
from hpcs_b200 import loss as L
from oracle import hpcs_oracle as O

REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "hpcs"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")


@pytest.fixture(scope="module")
def refmods():
    stale = [m for m in sys.modules if (m == "hpcs" or m.startswith("hpcs.")) and "fake_reference" in str(getattr(sys.modules[m], "__file__", ""))]
    for m in stale:
        del sys.modules[m]
    from oracle import ref_stubs
    ref_stubs.install_models()
    import hpcs.loss.ultrametric_loss as ul
    import hpcs.loss.hierarchical_cosface_loss as hcl
    import hpcs.distances.cosine as cos
    return ul, hcl, cos


def test_cosine_similarity_class():
    gen = torch.Generator().manual_seed(0)
    q, r = torch.randn(40, 8, generator=gen), torch.randn(30, 8, generator=gen)
    sim = hb.CosineSimilarity()
    assert sim.is_inverted and sim.normalize_embeddings
    want = O.cosine_similarity_matrix(q.double())
    assert torch.allclose(sim(q).double(), want, atol=1e-6)
    qn, rn = torch.nn.functional.normalize(q), torch.nn.functional.normalize(r)
    assert torch.allclose(sim(q, r), 0.5 * (1 + qn @ rn.t()), atol=1e-6)
    assert torch.allclose(sim.pairwise_distance(qn[:30], rn), 0.5 * (1 + (qn[:30] * rn).sum(1)), atol=1e-6)
    assert torch.equal(sim.margin(torch.tensor(0.2), torch.tensor(0.7)), torch.tensor(0.7) - torch.tensor(0.2))
    assert sim.smallest_dist(torch.tensor([0.1, 0.9])) == 0.9


@needs_ref
def test_cosine_similarity_matches_reference_class(refmods):
    _, _, cos = refmods
    gen = torch.Generator().manual_seed(1)
    q, r = torch.randn(50, 16, generator=gen), torch.randn(20, 16, generator=gen)
    ours, theirs = hb.CosineSimilarity(), cos.CosineSimilarity()
    assert torch.equal(ours(q), theirs(q)) and torch.equal(ours(q, r), theirs(q, r))
    assert ours.is_inverted == theirs.is_inverted
    assert torch.equal(ours.pairwise_distance(q[:20], r), theirs.pairwise_distance(q[:20], r))


@needs_ref
def test_cosface_and_logits_match_reference(refmods):
    """Same W -> same loss, same get_logits (ultrametric_loss.py:95-112), same state_dict key and shape
    (``loss_cosface.W`` [embedding_size, num_classes]: what reference checkpoints hold)."""
    ul, _, _ = refmods
    torch.manual_seed(0)
    theirs = ul.MetricHyperbolicLoss(num_class=7, embedding_size=5, cosface=True, miner=True, scale=torch.tensor([1e-3]))
    ours = hb.MetricHyperbolicLoss(num_class=7, embedding_size=5, cosface=True, miner=True, scale=torch.tensor([1e-3]))
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
    ours.load_state_dict(theirs.state_dict())
    x = (torch.randn(64, 5) * 0.2).requires_grad_(True)
    y = torch.randint(0, 7, (64,))
    assert torch.allclose(ours.get_logits(x, y), theirs.get_logits(x, y), atol=1e-6)
    lc = theirs.loss_cosface                            # the oracle's restatement (what the GPU kernel is tested against)
    assert torch.allclose(O.cosface_logits(x, lc.W, y, lc.margin, lc.scale), theirs.get_logits(x, y), atol=1e-6)
    lo, lt = ours.loss_cosface(x, y), theirs.loss_cosface(x, y)
    assert torch.allclose(lo, lt, atol=1e-6)
    go, gt = torch.autograd.grad(lo, x)[0], torch.autograd.grad(lt, x)[0]
    assert torch.allclose(go, gt, atol=1e-6)


@needs_ref
def test_hierarchical_cosface_matches_reference_class(refmods):
    ul, hcl, _ = refmods
    hier = [[[0, 1, 2], [3, 4], [5]], [[0], [1, 2], [3], [4, 5]], [[0], [1], [2], [3], [4], [5]]]
    torch.manual_seed(2)
    theirs = hcl.HierarchicalCosFaceLoss(num_classes=6, embedding_size=4, margin=0.35, scale=2, hierarchy_list=hier)
    ours = hb.HierarchicalCosFaceLoss(num_classes=6, embedding_size=4, margin=0.35, scale=2, hierarchy_list=hier)
    ours.load_state_dict(theirs.state_dict())
    x = (torch.randn(80, 4) * 0.3).requires_grad_(True)
    y = torch.randint(0, 6, (80,))
    lo, lt = ours(x, y), theirs(x, y)
    assert torch.allclose(lo, lt, rtol=1e-6, atol=1e-6)
    (go, gwo), (gt, gwt) = torch.autograd.grad(lo, (x, ours.W)), torch.autograd.grad(lt, (x, theirs.W))
    assert torch.allclose(go, gt, atol=1e-6) and torch.allclose(gwo, gwt, atol=1e-6)
    # the module that owns it keeps the reference's constructor and attribute names
    m = hb.HierarchicalMetricHyperbolicLoss(num_class=6, embedding_size=4, miner=True, hierarchy_list=hier)
    r = ul.HierarchicalMetricHyperbolicLoss(num_class=6, embedding_size=4, miner=True, hierarchy_list=hier)
    assert isinstance(m, hb.MetricHyperbolicLoss) and isinstance(m.loss_cosface, hb.HierarchicalCosFaceLoss)
    assert set(m.state_dict()) == set(r.state_dict())
    for name in ("margin", "t_per_anchor", "fraction", "temperature", "anneal_factor", "num_class", "embedding_size",
                 "cosface", "miner", "hierarchy_list"):
        assert getattr(m, name) == getattr(r, name), name


@needs_ref
def test_triplet_margin_branch_matches_reference(refmods):
    """cosface=False: relu(sim(a,n) - sim(a,p) + margin), mean over violating triplets -- the reference's own
    TripletMarginLoss on the same mined triplets."""
    ul, _, _ = refmods
    theirs = ul.MetricHyperbolicLoss(num_class=4, embedding_size=6, cosface=False, miner=True, margin=0.3,
                                     scale=torch.tensor([1e-3]), t_per_anchor=4, fraction=0.0)
    gen = torch.Generator().manual_seed(4)
    x = (torch.randn(90, 6, generator=gen) * 0.2).requires_grad_(True)
    y = torch.randint(0, 4, (90,), generator=gen)
    torch.manual_seed(11)
    trip = theirs.triplet_miner(x, y)                      # reference miner: semihard, margin 0.3
    assert trip[0].numel() > 0
    want = theirs.loss_triplet(x, y, trip)
    got = L.triplet_margin_loss(x, *trip, 0.3)
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    assert torch.allclose(torch.autograd.grad(got, x)[0], torch.autograd.grad(want, x)[0], atol=1e-6)
    empty = torch.empty(0, dtype=torch.long)
    assert L.triplet_margin_loss(x, empty, empty, empty, 0.3).item() == 0.0


def test_get_triplets_and_anneal():
    mod = hb.MetricHyperbolicLoss(t_per_anchor=3, temperature=0.1, anneal_factor=0.5, miner=False)
    torch.manual_seed(5)
    a, p, n = mod.get_triplets(12)
    assert a.numel() == p.numel() == n.numel() > 0
    assert (a < p).all() and (n != a).all() and (n != p).all() and int(n.max()) < 12
    if HAVE_REF:
        from oracle import ref_stubs
        ref_stubs.install_models()
        import hpcs.loss.ultrametric_loss as ul
        ref = ul.MetricHyperbolicLoss(t_per_anchor=3, miner=False)
        torch.manual_seed(5)
        ra, rp, rn = ref.get_triplets(12)
        assert torch.equal(a, ra) and torch.equal(p, rp) and torch.equal(n, rn)
    assert mod.anneal_temperature() == pytest.approx(0.05)   # the reference's raises (clamp on a float); documented
    mod.anneal_factor = 5.0
    assert mod.anneal_temperature() == pytest.approx(0.05)   # factor clamped to [0.2, 1]


def test_filter_mode_names():
    """'all' is the miner's constructor default and means 'margin test only' (triplet_margin_miner.py:27-32), not
    'keep everything' (ADVICE round 1)."""
    assert L.FILTER_MODES["all"] == 4 and L.FILTER_MODES["none"] == 0
    assert L.FILTER_MODES["easy"] == 1 and L.FILTER_MODES["semihard"] == 2 and L.FILTER_MODES["hard"] == 3
    miner = hb.RandomTripletMarginMiner(t_per_anchor=2, fraction=0.0)
    assert miner.type_of_triplets == "all"
