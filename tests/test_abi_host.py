"""CPU checks of the boundary: the shared library builds for sm_100a, loads, exports every symbol
that include/hpcs_b200.h declares, and the product refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hpcs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hpcs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hpcs_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/hpcs_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names            # the ctypes table mirrors the header one to one
    assert lib.hpcs_abi_version() == _lib.ABI_VERSION == 4
    assert _lib.launch_count() == 0 or torch.cuda.is_available()


def test_library_is_sm100a_only():
    from hpcs_b200 import build
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_workspace_queries_need_no_gpu():
    from hpcs_b200 import _lib
    lib = _lib.load()
    assert lib.hpcs_knn_workspace_bytes(32, 3, 1024, 20) >= 32 * 1024 * 4
    assert lib.hpcs_edge_feat_bwd_workspace_bytes(32, 1024, 20) >= 32 * 1024 * 20 * 8
    assert lib.hpcs_hyp_triplet_workspace_bytes(32768, 32) >= 2 * 32768 * 32 * 4
    assert lib.hpcs_linkage_workspace_bytes(2, 1024, 32, 0) >= 2 * 1024 * 1024 * 8
    assert lib.hpcs_hyp_triplet_workspace_bytes(10, 500) == 0       # unsupported dim -> 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_gpu():
    import hpcs_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hpcs_b200.knn(torch.randn(1, 3, 64), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hpcs_b200.get_graph_feature(torch.randn(1, 1, 3, 64), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hpcs_b200.hyp_lca(torch.rand(4, 8) * 0.1, torch.rand(4, 8) * 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hpcs_b200.decode_linkage(torch.randn(16, 8), torch.tensor([1e-3]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hpcs_b200.hyp_triplet_loss(torch.randn(16, 8), (torch.zeros(1), torch.zeros(1), torch.zeros(1)),
                                   torch.tensor([1e-3]), 0.05)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hpcs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+scipy\b", src, flags=re.M), f   # no CPU decoder either


def test_host_sampler_matches_oracle_sampler():
    from hpcs_b200.loss import get_balanced_random_triplet_indices
    from oracle import hpcs_oracle as O
    labels = torch.randint(0, 7, (500,), generator=torch.Generator().manual_seed(0))
    for frac in (0.0, 1.2):
        torch.manual_seed(42)
        got = get_balanced_random_triplet_indices(labels, t_per_anchor=5, fraction=frac)
        torch.manual_seed(42)
        want = O.sample_triplets(labels, 5, frac)
        for g_, w_ in zip(got, want):
            assert torch.equal(g_, w_)


def test_host_sampler_matches_reference_golden(golden):
    from hpcs_b200.loss import get_balanced_random_triplet_indices
    g = golden("compute_hyp")
    labels = torch.from_numpy(g["labels"]).long()
    for frac in (0.0, 1.2):
        torch.manual_seed(1234)
        a, p, n = get_balanced_random_triplet_indices(labels, t_per_anchor=7, fraction=frac)
        assert torch.equal(a, torch.from_numpy(g[f"f{frac}_a"]).long())
        assert torch.equal(p, torch.from_numpy(g[f"f{frac}_p"]).long())
        assert torch.equal(n, torch.from_numpy(g[f"f{frac}_n"]).long())
