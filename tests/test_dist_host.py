"""world_size-2 gloo tests (CPU) of the data-parallel helpers: cloud sharding is a partition, the
gradient all-reduce averages, and gathering per-rank dendrograms on rank 0 restores batch order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hpcs_b200 import dist as hdist


def test_shard_range_partitions():
    for total in (0, 1, 7, 32, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [hdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = hdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    clouds = torch.arange(10 * 4, dtype=torch.float32).view(10, 4)
    mine = hdist.shard_clouds(clouds, r, w)
    grads = [torch.full((3,), float(r + 1)), torch.full((2, 2), float(10 * (r + 1)))]
    hdist.allreduce_mean_(grads)
    gathered = hdist.gather_to_rank0(mine * 2)
    if r == 0:
        out["grads"] = [g.clone() for g in grads]
        out["gathered"] = torch.cat(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_gather():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.equal(out["grads"][0], torch.full((3,), 1.5))
    assert torch.equal(out["grads"][1], torch.full((2, 2), 15.0))
    assert torch.equal(out["gathered"], torch.arange(40, dtype=torch.float32).view(10, 4) * 2)
