"""BASELINE.json configs[3]: kNN-graph + hyperbolic-loss microbench sweep (N = 1024..16384 points, k = 10..40, batch 1..256),
one markdown table on stdout.  Every cell is also a correctness spot check: kNN indices against the all-FFMA kernels
(two independent code paths, both bit-exact against the oracle in the tests), the edge gradient of one cloud against a
float64 PyTorch scatter, the loss for finiteness (its parity against the fp64 oracle is what tests/ cover).
  python tools/sweep_bench.py [--quick]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpcs_b200 as hb  # noqa: E402
from hpcs_b200 import graph as hgraph  # noqa: E402


def timed(fn, reps=10, warm=3):
    """Average duration of fn in microseconds: captured into a CUDA graph and replayed (no Python / launch overhead in the
    number, like bench.py's per-op timings); eager as a fallback."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    run = fn
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        run = graph.replay
    except Exception:                                   # e.g. an op that cannot be captured: time it eagerly
        torch.cuda.synchronize()
    run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3          # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    grid = [(32, 1024, 20), (1, 1024, 20), (8, 1024, 10), (256, 1024, 20), (32, 1024, 40), (32, 2048, 20), (8, 4096, 20),
            (8, 8192, 10), (4, 16384, 20)]
    if args.quick:
        grid = grid[:3]
    print("| B | N | k | kNN D=3 µs | kNN D=63 µs | path | edge fwd C=21 µs (TB/s) | edge bwd C=21 µs (TB/s) | fused EdgeConv layer (C=21, two convs) fwd eval / fwd+bwd train, µs | loss fwd+bwd µs (50 triplets/pt) | checks |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    from hpcs_b200.edgeconv import edgeconv

    class Conv(torch.nn.Module):
        def __init__(self, cin):
            super().__init__()
            self.negative_slope = 0.2
            self.map_to_feat = torch.nn.Linear(cin, 21, bias=False)
            self.map_to_dir = torch.nn.Linear(cin, 21, bias=False)
            self.batchnorm = torch.nn.Module()
            self.batchnorm.bn = torch.nn.BatchNorm2d(21)
    convs = [Conv(42).to(dev), Conv(21).to(dev)]
    for B, N, k in grid:
        x3 = torch.randn(B, 3, N, device=dev, generator=gen)
        x63 = torch.randn(B, 63, N, device=dev, generator=gen)
        t3 = timed(lambda: hb.knn(x3, k))
        t63 = timed(lambda: hb.knn(x63, k))
        ok = []
        i3, i63 = hb.knn(x3, k), hb.knn(x63, k)
        ok.append("knn3" if torch.equal(i3, hb.knn(x3, k, method="ffma")) else "KNN3-MISMATCH")
        ok.append("knn63" if torch.equal(i63, hb.knn(x63, k, method="ffma")) else "KNN63-MISMATCH")
        path = "tcgen05" if N <= 16384 and k <= 48 else "ffma"
        # edge features (C = 21): skip shapes whose output would not fit comfortably
        C = 21
        out_bytes = B * 2 * C * 3 * N * k * 4
        if out_bytes <= 24e9:
            x = x63.view(B, C, 3, N)
            g = torch.randn(B, 2 * C, 3, N, k, device=dev, generator=gen)
            tf = timed(lambda: hgraph.edge_features_forward(x, i63), reps=5)
            tb = timed(lambda: hgraph.edge_features_backward(g, x, i63), reps=5)
            fast = bool(hb._lib.load().hpcs_edge_feat_bwd_is_fast(g.data_ptr(), N, k, 0))
            byt = out_bytes + B * 3 * C * N * 4 + B * N * k * 8
            # gradient check on one cloud in float64
            gx = hgraph.edge_features_backward(g, x, i63)[0].double()
            g0, idx0 = g[0].double(), i63[0]
            ref = g0[C:].sum(-1) - g0[:C].sum(-1)
            ref = ref.reshape(C * 3, N)
            ref.index_add_(1, idx0.reshape(-1), g0[:C].reshape(C * 3, N * k))
            err = (gx.reshape(C * 3, N) - ref).abs().max().item() / ref.abs().max().item()
            ok.append("edge" if err < 1e-5 else f"EDGE-ERR {err:.1e}")
            how = "" if fast else ", scatter path"
            edge = f"{tf:.0f} ({byt / tf / 1e6:.2f}) | {tb:.0f} ({byt / tb / 1e6:.2f}{how})"
            del g, gx
        else:
            edge = "— | —"
        # the fused layer covers every shape (row gathers from L2, no per-cloud shared-memory structure); eager launches
        def eager(fn, reps=3):
            fn(); torch.cuda.synchronize()
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            for _ in range(reps):
                fn()
            e_.record(); torch.cuda.synchronize()
            return s_.elapsed_time(e_) / reps * 1e3
        xl = x63.view(B, 21, 3, N)
        gl = torch.randn(B, 21, 3, N, device=dev, generator=gen)
        for c_ in convs:
            c_.eval()
        with torch.no_grad():
            tfe = eager(lambda: edgeconv(xl, k, convs[0], convs[1], idx=i63))
        for c_ in convs:
            c_.train()

        def layer_step():
            xr = xl.detach().requires_grad_(True)
            y = edgeconv(xr, k, convs[0], convs[1], idx=i63)
            torch.autograd.grad((y * gl).sum(), [xr] + [p_ for c_ in convs for p_ in c_.parameters()])
        tft = eager(layer_step)
        fusedcol = f"{tfe:.0f} / {tft:.0f}"
        # loss on n = B*N points (capped), 50 triplets per anchor, device sampler
        n = min(B * N, 65536)
        emb = torch.randn(n, 32, device=dev, generator=gen)
        emb = torch.tanh(emb.norm(dim=-1, keepdim=True)) * emb / emb.norm(dim=-1, keepdim=True)
        labels = torch.randint(0, 8, (n,), generator=torch.Generator().manual_seed(1))
        order, seg, T0 = hb.triplet_plan(labels, 50, 0.0)
        trip = hb.sample_triplets_device(None, seed=3, plan=(order.to(dev), seg.to(dev), T0))
        sc = torch.tensor([1e-3], device=dev)

        def loss_step():
            e = emb.detach().requires_grad_(True)
            s_ = sc.detach().requires_grad_(True)
            l = hb.hyp_triplet_loss(e, trip, s_, 0.05, "easy", 0.0)
            torch.autograd.grad(l, (e, s_))
            return l
        tl = timed(loss_step, reps=5)
        ok.append("loss" if torch.isfinite(loss_step()) else "LOSS-NAN")
        print(f"| {B} | {N} | {k} | {t3:.0f} | {t63:.0f} | {path} | {edge} | {fusedcol} | {tl:.0f} ({T0 / tl:.0f} M triplets/s, n={n}) | {' '.join(ok)} |", flush=True)
        del x3, x63, emb
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
