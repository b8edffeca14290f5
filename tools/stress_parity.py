"""Randomised bit-exactness stress of the index / dendrogram kernels against the oracle (test infrastructure, like tests/):
  python tools/stress_parity.py [--seconds 120] [--seed 0]
kNN (every path: D=3 FFMA, tensor-core + re-rank, second chance, exact redo) vs oracle.knn_canonical; linkage (single and
complete, parallel rounds and serial) vs scipy.  Data kinds are chosen to hit the rare paths: clustered, offset, duplicated,
low-rank and grid-quantised (exact ties) clouds.  Prints one line per failure and a summary; exit code 1 on any mismatch."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpcs_b200 as hb  # noqa: E402
from hpcs_b200 import decode  # noqa: E402
from hpcs_b200.hyperbolic import normalize_project  # noqa: E402
from oracle import hpcs_oracle as O  # noqa: E402


def make_cloud(gen, B, D, N, kind):
    x = torch.randn(B, D, N, generator=gen)
    if kind == "clustered":
        c = torch.randn(B, D, 8, generator=gen) * 4
        x = c[:, :, torch.randint(0, 8, (N,), generator=gen)] + 0.05 * x
    elif kind == "offset":
        x = x * 0.1 + 7.0
    elif kind == "dup":
        src = torch.randint(0, N, (N,), generator=gen)
        m = torch.rand(N, generator=gen) < 0.3
        x[:, :, m] = x[:, :, src[m]]
    elif kind == "lowrank":
        r = max(1, D // 8)
        x = torch.randn(B, D, r, generator=gen) @ torch.randn(B, r, N, generator=gen)
    elif kind == "grid":
        x = torch.round(x * 2) / 2
    elif kind == "scaled":
        x = x * torch.exp(torch.randn(B, D, 1, generator=gen) * 2)
    return x.contiguous()


KINDS = ("gauss", "clustered", "offset", "dup", "lowrank", "grid", "scaled")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", default="knn,linkage")
    args = ap.parse_args()
    gen = torch.Generator().manual_seed(args.seed)
    rng = np.random.default_rng(args.seed)
    from scipy.cluster.hierarchy import linkage
    t_end = time.time() + args.seconds
    n_knn = n_link = bad = 0
    paths = {"fallback_rows": 0, "second_chance_rows": 0}
    while time.time() < t_end:
        # ---- kNN ----
        do_knn, do_link = "knn" in args.only, "linkage" in args.only
        D = int(rng.choice([3, 3, 16, 24, 33, 48, 63, 63, 63]))
        N = int(rng.choice([40, 64, 100, 256, 500, 768, 1024, 1024, 1500, 2048]))
        B = int(rng.integers(1, 5))
        k = int(rng.integers(2, min(24, N - 1) + 1))
        kind = KINDS[int(rng.integers(0, len(KINDS)))]
        x = make_cloud(gen, B, D, N, kind)
        st = {}
        got = hb.knn(x.cuda(), k, stats=st).cpu()
        want = O.knn_canonical(x, k)
        for key in paths:
            paths[key] += int(st.get(key, 0))
        n_knn += 1
        if not torch.equal(got, want):
            bad += 1
            print(f"KNN MISMATCH B={B} D={D} N={N} k={k} kind={kind} rows={(got != want).any(-1).sum().item()}", flush=True)
        # ---- linkage ----
        N = int(rng.choice([2, 3, 5, 17, 64, 130, 257, 600, 1024]))
        D = int(rng.choice([2, 4, 16, 32, 33, 64]))
        B = int(rng.integers(1, 4))
        kind = KINDS[int(rng.integers(0, len(KINDS)))]
        e = make_cloud(gen, B, D, N, kind).transpose(1, 2).contiguous()            # [B, N, D] embeddings
        e = torch.where(e.abs().sum(-1, keepdim=True) == 0, torch.ones_like(e), e)   # cosine distance of a zero vector is NaN
        e = e * 0.3 / e.norm(dim=-1, keepdim=True).clamp_min(1e-3) * torch.rand(B, N, 1, generator=gen)
        scale = torch.tensor([1.0])
        leaves = normalize_project(e.cuda(), scale.cuda()).cpu().numpy().astype(np.float64)
        for method in ("single", "complete"):
            for force in (None, "serial", "rounds") if method == "complete" else (None,):
                if force:
                    os.environ["HPCS_COMPLETE_LINKAGE"] = force
                else:
                    os.environ.pop("HPCS_COMPLETE_LINKAGE", None)
                Z = decode.decode_linkage_batch(e.cuda(), scale.cuda(), method=method).cpu().numpy()
                for b in range(B):
                    # same definition as tests/test_gpu_parity.py: scipy on the leaves the decoder used (the fp32 normalisation
                    # itself is compared with the oracle to 1e-6 there; a last-bit difference in a leaf moves every height)
                    want = linkage(leaves[b], method=method, metric="cosine")
                    n_link += 1
                    if not np.array_equal(Z[b], want):
                        bad += 1
                        print(f"LINKAGE MISMATCH method={method} force={force} B={B} N={N} D={D} kind={kind} cloud={b}", flush=True)
        os.environ.pop("HPCS_COMPLETE_LINKAGE", None)
    print(f"stress: {n_knn} kNN cases ({paths}), {n_link} dendrograms, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
