"""Where one whole-model training step spends its GPU time (the `model_step` record of bench.py, same construction):
  python tools/model_step_profile.py [--B 32]
torch.profiler over 3 steps after warm-up; kernels grouped into: this library's kernels, the dense tail / CosFace / optimizer
(PyTorch library kernels)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=bench.B_PER_GPU)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    fake = os.path.join(ROOT, "tests", "fake_reference")
    sys.path.insert(0, fake)
    import hpcs.models, hpcs.nn.dgcnn, hpcs.nn.pointnet, hpcs.nn.hyperbolic   # noqa: F401,E401
    from hpcs_b200 import patch
    patch.install(strict=True)
    from hpcs.models import ShapeNetHypHC
    from hpcs.nn.dgcnn import VN_DGCNN_partseg
    from hpcs.nn.hyperbolic import ExpMap
    VN_DGCNN_partseg.tail_width = 1024 // 3
    B, N = args.B, bench.N_PTS
    torch.manual_seed(0)
    model = ShapeNetHypHC(nn_feat=VN_DGCNN_partseg(3, bench.D_EMB, bench.K_NN, 0.5, "mean", 16), nn_emb=ExpMap(), euclidean_size=bench.D_EMB,
                          hyp_size=bench.D_EMB, num_class=50, t_per_anchor=bench.T_PER_ANCHOR, fraction=0.0,
                          temperature=bench.TEMPERATURE, miner=True).to(dev).train()
    opt = torch.optim.RAdam(model.parameters(), lr=1e-3)
    host = bench.synth_inputs(B, 1234)
    pts = host["pts"].view(B, 3, N).transpose(1, 2).contiguous()
    targets = host["labels"].view(B, N)
    label = torch.randint(0, 16, (B, 1), generator=torch.Generator().manual_seed(5))

    def step():
        losses, _ = model.forward((pts, label, targets), testing=False)
        total = losses["loss_metric"] + losses["loss_hyp"]
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t and e.device_type.name == "CUDA":
            rows.append((t / 3.0, e.count // 3, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    ours = sum(r[0] for r in rows if "hpcs::" in r[2])
    print(f"GPU kernel time per step: {total / 1e3:.2f} ms; hpcs_b200 kernels {ours / 1e3:.2f} ms ({100 * ours / total:.1f} %), "
          f"PyTorch library kernels {(total - ours) / 1e3:.2f} ms")
    for t, n, k in rows[:28]:
        print(f"{t / 1e3:9.3f} ms x{n:4d}  {k[:150]}")


if __name__ == "__main__":
    main()
