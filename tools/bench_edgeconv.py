#!/usr/bin/env python
"""Time the fused EdgeConv layers (row f-1) against the unfused composition on the GPU.

  python tools/bench_edgeconv.py [--B 32] [--N 1024] [--k 20] [--reps 10] [--unfused]

Per layer shape of VN_DGCNN_partseg (conv1+conv2 on coordinates, conv3+conv4 and conv5 on 21 vector channels): fused
forward (training mode: statistics passes included), fused forward+backward, each kernel launch separately (CUDA events
around repeated calls), and -- with --unfused -- hpcs_b200.get_graph_feature + the same VN arithmetic as plain PyTorch
ops over the [B,2C,3,N,k] tensor (what the reference's modules execute), forward+backward, with peak memory."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--unfused", action="store_true")
    args = ap.parse_args()
    import hpcs_b200 as hb
    from hpcs_b200.edgeconv import edgeconv
    from test_gpu_edgeconv import VNConv, param_list
    from oracle import hpcs_oracle as O
    B, N, k = args.B, args.N, args.k
    out = {"B": B, "N": N, "k": k}
    torch.manual_seed(0)
    for tag, C, two in (("layer1_c1_2conv", 1, True), ("layer2_c21_2conv", 21, True), ("layer3_c21_1conv", 21, False)):
        convs = [VNConv(2 * C).cuda().train()] + ([VNConv(21).cuda().train()] if two else [])
        x = torch.randn(B, C, 3, N, device="cuda")
        gout = torch.randn(B, 21, 3, N, device="cuda")
        idx = hb.knn(x.view(B, 3 * C, N), k)
        rec = {}

        def fwd():
            with torch.no_grad():
                return edgeconv(x, k, convs[0], convs[1] if two else None, idx=idx)

        def fwd_bwd():
            xr = x.detach().requires_grad_(True)
            y = edgeconv(xr, k, convs[0], convs[1] if two else None, idx=idx)
            torch.autograd.grad((y * gout).sum(), [xr] + param_list(convs))
        rec["fused_fwd_train_ms"] = round(timed(fwd, args.reps), 4)
        rec["fused_fwd_bwd_ms"] = round(timed(fwd_bwd, args.reps), 4)
        for c in convs:
            c.eval()
        rec["fused_fwd_eval_ms"] = round(timed(fwd, args.reps), 4)
        for c in convs:
            c.train()
        torch.cuda.reset_peak_memory_stats()
        fwd_bwd()
        rec["fused_peak_mem_mb"] = round(torch.cuda.max_memory_allocated() / 1e6, 1)
        if args.unfused:
            def unfused():
                xr = x.detach().requires_grad_(True)
                e = hb.get_graph_feature(xr, k, idx=idx)
                for c in convs:
                    bn = c.batchnorm.bn
                    e = O.vn_linear_leaky_relu(e, c.map_to_feat.weight, c.map_to_dir.weight, bn.weight, bn.bias, None, None, True)
                y = e.mean(dim=-1)
                torch.autograd.grad((y * gout).sum(), [xr] + param_list(convs))
            torch.cuda.reset_peak_memory_stats()
            rec["unfused_fwd_bwd_ms"] = round(timed(unfused, max(2, args.reps // 3)), 4)
            rec["unfused_peak_mem_mb"] = round(torch.cuda.max_memory_allocated() / 1e6, 1)
            rec["speedup_fwd_bwd"] = round(rec["unfused_fwd_bwd_ms"] / rec["fused_fwd_bwd_ms"], 2)
        out[tag] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
