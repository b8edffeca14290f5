"""Time the edge-feature backward at the bench shape (CUDA-graph replay) and print the per-phase cycle split of the
persistent gather.  The split needs a profiling build:  add "-DHPCS_BWD_PROFILE" to NVCC_FLAGS in hpcs_b200/build.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpcs_b200 as hb  # noqa: E402
from hpcs_b200 import graph as hgraph  # noqa: E402


def graph_time(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    B, C, N, K = 32, 21, 1024, 20
    x = torch.randn(B, C, 3, N, device=dev)
    idx = hb.knn(x.view(B, 3 * C, N), K)
    g = torch.randn(B, 2 * C, 3, N, K, device=dev)
    bytes_ = g.numel() * 4 + idx.numel() * 8 + x.numel() * 4
    prof = torch.zeros(8 * 320, dtype=torch.int64, device=dev)
    os.environ["HPCS_BWD_PROF_PTR"] = hex(prof.data_ptr())
    hgraph.edge_features_backward(g, x, idx)
    torch.cuda.synchronize()
    del os.environ["HPCS_BWD_PROF_PTR"]
    pr = prof.view(320, 8).cpu().double()
    pr = pr[pr.sum(1) > 0]
    names = ["meta-tail", "wait", "centre", "gather", "combine", "issue", "meta-loads", "meta-sync1"]
    tot = pr.sum(1).mean().item()
    print("phase cycles per CTA (mean over %d CTAs), total %.0f:" % (pr.shape[0], tot))
    for i, nme in enumerate(names):
        print(f"   {nme:8s} {pr[:, i].mean().item():10.0f}  {100 * pr[:, i].mean().item() / tot:5.1f}%   max {pr[:, i].max().item():10.0f}")
    us = graph_time(lambda: hgraph.edge_features_backward(g, x, idx))
    print(f"backward (reverse-graph build + gather): {us:8.1f} us   {bytes_ / us / 1e3:7.1f} GB/s", flush=True)
    rev = hgraph.build_reverse_graph(idx, overlap=False)
    us = graph_time(lambda: hgraph.edge_features_backward(g, x, idx, prebuilt=rev))
    print(f"gather alone (prebuilt reverse graph):   {us:8.1f} us   {bytes_ / us / 1e3:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
