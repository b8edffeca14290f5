"""Data-parallel helpers: one process per GPU, clouds sharded over ranks.

Point clouds are independent, so kNN / edge features / decode shard along the batch dimension with
no data-path collective; the loss couples only the points of the local shard, which is the
reference's DDP semantics.  The one exchange step of a training iteration
is the all-reduce (mean) of parameter gradients over NCCL.
"""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group when
    WORLD_SIZE > 1.  Rendezvous uses MASTER_ADDR/MASTER_PORT as given (127.0.0.1 on one node)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        # The one collective of this path is a 5 MB gradient bucket overlapped with persistent one-CTA-per-SM kernels: every
        # channel NCCL opens is a CTA on an SM those kernels are waiting for (144 gather CTAs on 148 SMs), but too few channels
        # make the all-reduce longer than the window it hides in.  Measured per batch, in-graph, B200 (profiles/r02_bench_*gpu*):
        #   2 GPUs (ring):  0.825 ms with NCCL's defaults, 0.806 with NCCL_MAX_NCHANNELS=2, 0.841 with 8;
        #   4 GPUs:         0.821 ms with the defaults, 0.891 with NCCL_MAX_NCHANNELS=2, 0.994 with 4, 0.832 with NVLS_NCHANNELS=2;
        #   8 GPUs (NVLS):  0.858 ms with the defaults (24 NVLS channels; NCCL_MAX_NCHANNELS makes no difference), 0.839 with
        #                   NCCL_NVLS_NCHANNELS=2; ring without NVLS 1.08-1.38 ms.
        # The all-reduce alone takes 56 us at 8 GPUs whatever the setting (tools/allreduce_probe.py).  So: two ring channels
        # at 2 GPUs, NCCL's own choice at 4, two NVLS channels at 8 and beyond; an explicit setting in the environment wins.
        if world == 2:
            os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
        elif world >= 8:
            os.environ.setdefault("NCCL_NVLS_NCHANNELS", "2")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``total`` units owned by ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_clouds(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_range(t.shape[0], rank, world)
    return t[b:e]


def allreduce_mean_(tensors: Iterable[torch.Tensor]) -> None:
    """In-place mean all-reduce of gradient tensors, flattened into one bucket (a few MB: the
    collective is latency-bound on NVLink, so one launch beats many)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    ts = [t for t in tensors if t is not None]
    if not ts:
        return
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for t in ts:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def gather_to_rank0(t: torch.Tensor):
    """Collect per-rank results (e.g. dendrograms) on rank 0; returns a list there, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [t]
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(t.cpu(), out, dst=0)
    return out
