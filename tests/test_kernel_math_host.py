"""CPU check of the scalar formulas the CUDA kernels use (hpcs_b200/csrc/hyp_math.cuh compiled
with g++), against the oracle's fp64 evaluation of the reference formulas."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import hpcs_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(ROOT, "tests", "csrc", "hyp_math_check.cpp")
    out_dir = os.path.join(ROOT, "tests", "csrc", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libhyp_math_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    lib = ctypes.CDLL(out)
    f, d, i = ctypes.c_float, ctypes.c_double, ctypes.c_int
    lib.check_lca_equal_f32.argtypes = [f, f, ctypes.POINTER(f)]
    lib.check_lca_equal_f64.argtypes = [d, d, ctypes.POINTER(d)]
    lib.check_triplet_f32.argtypes = [f, f, f, f, f, i, f, ctypes.POINTER(f)]
    lib.check_triplet_f64.argtypes = [d, d, d, d, d, i, d, ctypes.POINTER(d)]
    lib.check_lca_general.argtypes = [d, d, d, ctypes.POINTER(d)]
    return lib


def _pair(cos, s, D=8):
    """Two fp64 vectors of norm s with the given cosine."""
    a = torch.zeros(D, dtype=torch.float64); a[0] = s
    b = torch.zeros(D, dtype=torch.float64); b[0] = s * cos; b[1] = s * np.sqrt(max(0.0, 1 - cos * cos))
    return a, b


@pytest.mark.parametrize("s", [1e-3, 1e-2, 0.1, 0.5, 0.9])
def test_equal_radius_closed_form(hm, s):
    rng = np.random.default_rng(0)
    for cos in np.concatenate([rng.uniform(-0.95, 0.999, 40), [0.0, 0.9999, -0.99]]):
        a, b = _pair(float(cos), s)
        a.requires_grad_(True); b.requires_grad_(True)
        ref = O.hyp_lca(a[None], b[None], return_coord=False)[0, 0]
        o64 = (ctypes.c_double * 3)()
        hm.check_lca_equal_f64(float(cos), s, o64)
        assert abs(o64[0] - ref.item()) <= 1e-8 * abs(ref.item())
        o32 = (ctypes.c_float * 3)()
        hm.check_lca_equal_f32(float(cos), s, o32)
        assert abs(o32[0] - ref.item()) <= 2e-6 * abs(ref.item())
        # derivatives by central differences of the fp64 closed form
        h = 1e-6
        p, m = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
        hm.check_lca_equal_f64(float(cos) + h, s, p); hm.check_lca_equal_f64(float(cos) - h, s, m)
        assert abs(o64[1] - (p[0] - m[0]) / (2 * h)) <= 1e-5 * abs(o64[1]) + 1e-12
        hs = s * 1e-6
        hm.check_lca_equal_f64(float(cos), s + hs, p); hm.check_lca_equal_f64(float(cos), s - hs, m)
        assert abs(o64[2] - (p[0] - m[0]) / (2 * hs)) <= 1e-5 * abs(o64[2]) + 1e-12
        assert abs(o32[1] - o64[1]) <= 1e-4 * abs(o64[1]) + 1e-9
        assert abs(o32[2] - o64[2]) <= 1e-4 * abs(o64[2]) + 1e-9


@pytest.mark.parametrize("s,temp", [(1e-3, 0.05), (0.1, 0.05), (0.5, 0.1), (0.9, 0.05)])
def test_triplet_terms_match_oracle_autograd(hm, s, temp):
    """total and its derivatives wrt the three cosines and s, vs autograd through the oracle."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        c = torch.tensor(rng.uniform(-0.9, 0.98, 3), dtype=torch.float64, requires_grad=True)
        sv = torch.tensor(s, dtype=torch.float64, requires_grad=True)

        def lca(cosv):
            # build the pair inside autograd: a = s e0, b = s (cos e0 + sin e1)
            a = torch.stack([sv, sv * 0]).unsqueeze(0)
            b = torch.stack([sv * cosv, sv * torch.sqrt(1 - cosv * cosv)]).unsqueeze(0)
            return O.hyp_lca(a, b, return_coord=False)[0, 0]

        d = torch.stack([lca(c[0]), lca(c[1]), lca(c[2])])
        w = 0.5 * (1 + c)
        soft = torch.softmax(d / temp, dim=0)
        total = w.sum() - (w * soft).sum()
        gc, gs = torch.autograd.grad(total, (c, sv))
        o = (ctypes.c_double * 6)()
        hm.check_triplet_f64(c[0].item(), c[1].item(), c[2].item(), s, 1.0 / temp, 0, 0.0, o)
        assert abs(o[0] - total.item()) <= 1e-9 * abs(total.item())
        for q in range(3):
            assert abs(o[1 + q] - gc[q].item()) <= 1e-6 * abs(gc[q].item()) + 1e-10
        assert abs(o[4] - gs.item()) <= 1e-6 * abs(gs.item()) + 1e-10
        o32 = (ctypes.c_float * 6)()
        hm.check_triplet_f32(c[0].item(), c[1].item(), c[2].item(), s, 1.0 / temp, 0, 0.0, o32)
        assert abs(o32[0] - total.item()) <= 2e-6 * abs(total.item())
        for q in range(3):
            assert abs(o32[1 + q] - gc[q].item()) <= 1e-4 * abs(gc[q].item()) + 2e-5   # fp32: terms of size ~1/temperature cancel
        assert abs(o32[4] - gs.item()) <= 2e-4 * abs(gs.item()) + 2e-5


def test_filter_modes(hm):
    o = (ctypes.c_float * 6)()
    cases = [(1, 0.0, 0.5, 0.2, 1), (1, 0.0, 0.2, 0.5, 0), (2, 0.35, 0.5, 0.2, 1), (2, 0.35, 0.2, 0.5, 0),
             (2, 0.1, 0.9, 0.1, 0), (3, 0.35, 0.2, 0.5, 1), (3, 0.35, 0.5, 0.2, 0), (0, 0.0, -1.0, 1.0, 1)]
    for mode, margin, c_ap, c_an, want in cases:
        hm.check_triplet_f32(c_ap, c_an, 0.0, 0.5, 20.0, mode, margin, o)
        assert int(o[5]) == want, (mode, margin, c_ap, c_an)


def test_general_lca_duals(hm, golden):
    g = golden("hyp_lca")
    for tag in ("mixed", "s0.1", "s1e-3", "s0.9"):
        a = torch.from_numpy(g[tag + "_a"]).double()
        b = torch.from_numpy(g[tag + "_b"]).double()
        dist_ref = g[tag + "_dist"][:, 0]
        coord_ref = g[tag + "_coord"]
        ga_ref, gb_ref = g[tag + "_ga"], g[tag + "_gb"]
        for t in range(a.shape[0]):
            A, B, ab = (a[t] @ a[t]).item(), (b[t] @ b[t]).item(), (a[t] @ b[t]).item()
            o = (ctypes.c_double * 12)()
            hm.check_lca_general(A, B, ab, o)
            coord = o[0] * a[t].numpy() + o[1] * b[t].numpy()
            np.testing.assert_allclose(coord, coord_ref[t], rtol=1e-6, atol=1e-12)
            rc = min(o[2], 1 - 1e-5)
            np.testing.assert_allclose(np.log1p(2 * rc / (1 - rc)), dist_ref[t], rtol=1e-7)
            k = 2.0 / (1 - rc * rc)
            ga = k * (2 * o[9] * a[t].numpy() + o[11] * b[t].numpy())
            gb = k * (2 * o[10] * b[t].numpy() + o[11] * a[t].numpy())
            np.testing.assert_allclose(ga, ga_ref[t], rtol=2e-5, atol=1e-9)
            np.testing.assert_allclose(gb, gb_ref[t], rtol=2e-5, atol=1e-9)


def test_device_sampler_plan_matches_reference_sampler_structure():
    """triplet_segments (the host plan of the device sampler) must give the reference sampler's triplet count and its
    anchor sequence: labels ascending, members ascending, k_l repeats -- for fraction 0 and 1.2, with singleton labels."""
    import torch
    from hpcs_b200.loss import triplet_segments, get_balanced_random_triplet_indices
    gen = torch.Generator().manual_seed(3)
    for fraction, t in ((0.0, 7), (1.2, 5), (0.5, 3)):
        labels = torch.randint(0, 9, (400,), generator=gen)
        labels[labels == 4] = 5                       # label 4 absent
        labels[17] = 40                               # a singleton label: skipped by the reference
        torch.manual_seed(0)
        a, p, n = get_balanced_random_triplet_indices(labels, t_per_anchor=t, fraction=fraction)
        seg, T0 = triplet_segments(torch.bincount(labels), t, fraction)
        assert T0 == a.numel()
        order = torch.sort(labels, stable=True)[1]
        want = torch.cat([order[s:s + m].repeat_interleave(r) for s, m, r, _ in seg.t().tolist()])
        assert torch.equal(want, a)
        assert seg[3].tolist() == (torch.cumsum(seg[1] * seg[2], 0) - seg[1] * seg[2]).tolist()
