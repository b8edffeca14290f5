"""Time the gradient-bucket all-reduce of the step (5.2 MB fp32) on its own, max over ranks.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_probe.py [--floats F]
NCCL's algorithm / protocol / channel choices are read from the environment at init, so one setting = one launch."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hpcs_b200 import dist as hdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--floats", type=int, default=bench.GRAD_BUCKET_FLOATS)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    rank, world, local = hdist.init_from_env()
    dev = torch.device("cuda", local)
    buf = torch.randn(args.floats, device=dev)
    for _ in range(20):
        dist.all_reduce(buf)
        buf.mul_(1.0 / world)
    torch.cuda.synchronize()
    dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.reps):
        dist.all_reduce(buf)
    t1.record()
    torch.cuda.synchronize()
    us = torch.tensor([t0.elapsed_time(t1) / args.reps * 1e3], device=dev)
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    if rank == 0:
        env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
        print(f"all_reduce {args.floats * 4} B x {world} ranks: {us.item():.1f} us  {env}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
