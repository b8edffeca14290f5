"""Run each hot-path op a few times on the bench shapes (for ncu launch lists / captures).
  python tools/prof_ops.py [--ops knn3,knn63,edge,loss,sampler,decode] [--reps 3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import hpcs_b200 as hb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="knn3,knn63,edge,loss")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--time", action="store_true", help="print CUDA-event timings per op")
    ap.add_argument("--warm", type=int, default=2, help="untimed warm-up calls per op (0 for a compact ncu capture)")
    args = ap.parse_args()
    ops = args.ops.split(",")
    dev = torch.device("cuda:0")
    B, N, K, C = bench.B_PER_GPU, bench.N_PTS, bench.K_NN, bench.C_FEAT
    host = bench.synth_inputs(B, 0)
    d = {k: v.to(dev) for k, v in host.items()}

    def timeit(name, fn):
        for _ in range(args.warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        if args.time:
            print(f"{name:14s} {s.elapsed_time(e) / args.reps * 1e3:9.1f} us")

    if "knn3" in ops:
        timeit("knn_d3", lambda: hb.knn(d["pts"].view(B, 3, N), K))
    if "knn63" in ops:
        timeit("knn_d63", lambda: hb.knn(d["f1"].view(B, 3 * C, N), K))
    if "edge" in ops:
        x = d["f1"].clone().requires_grad_(True)
        idx = hb.knn(x.detach().view(B, 3 * C, N), K)
        g = torch.randn(B, 2 * C, 3, N, K, device=dev)
        y = [None]

        def fwd():
            y[0] = hb.get_graph_feature(x, K, idx=idx)
        timeit("edge_fwd_c21", fwd)
        timeit("edge_bwd_c21", lambda: torch.autograd.grad(y[0], x, g, retain_graph=True))
    if "loss" in ops:
        torch.manual_seed(1000)
        trip = tuple(t.to(dev) for t in hb.get_balanced_random_triplet_indices(host["labels"], t_per_anchor=bench.T_PER_ANCHOR, fraction=0.0))
        emb = d["emb"].clone().requires_grad_(True)
        sc = torch.tensor([bench.SCALE], device=dev, requires_grad=True)

        def loss():
            l = hb.hyp_triplet_loss(emb, trip, sc, bench.TEMPERATURE, "easy", 0.0)
            torch.autograd.grad(l, (emb, sc))
        timeit("hyp_loss", loss)
    if "sampler" in ops:
        order, seg, T0 = hb.triplet_plan(host["labels"], bench.T_PER_ANCHOR, 0.0)
        plan = (order.to(dev), seg.to(dev), T0)
        timeit("triplet_sampler", lambda: hb.sample_triplets_device(None, seed=1, plan=plan))
    if "decode" in ops:
        x = d["emb"].view(B, N, -1)
        sc = torch.tensor([bench.SCALE], device=dev)
        timeit("decode_single", lambda: hb.decode_linkage_batch(x, sc, "single"))
        timeit("decode_complete", lambda: hb.decode_linkage_batch(x, sc, "complete"))


if __name__ == "__main__":
    main()
