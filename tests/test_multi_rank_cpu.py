"""The N > 1 path of bench.py without GPUs: gloo, world_size 2.  Structures are sharded by cost across ranks with no
data-path collective; the only communication is the barrier + MAX-reduction of the timed region."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pesto_b200.sharding import lpt_partition, rank_shard
from pesto_b200.synth import interfaceome_sizes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, costs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = rank_shard(costs, rank, world)
    # every rank "processes" its own structures; timing is the max over ranks, work is the sum over ranks
    atoms = torch.tensor([float(sum(int(costs[i]) for i in mine))])
    ms = torch.tensor([10.0 * (rank + 1)])
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(atoms, op=dist.ReduceOp.SUM)
    ids = [None] * world
    dist.all_gather_object(ids, mine)
    if rank == 0:
        torch.save({"ms": ms.item(), "atoms": atoms.item(), "ids": ids}, os.path.join(out_dir, "r0.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction(tmp_path):
    costs = (interfaceome_sizes(200) * 8).tolist()          # BASELINE config 5 size distribution, atoms per structure
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), costs, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(os.path.join(tmp_path, "r0.pt"))
    assert r["ms"] == 20.0                                    # max over ranks
    assert r["atoms"] == float(sum(costs))                    # whole-job units = sum over ranks
    flat = sorted(i for shard in r["ids"] for i in shard)
    assert flat == list(range(len(costs)))                    # disjoint and complete
    loads = [sum(costs[i] for i in shard) for shard in r["ids"]]
    assert max(loads) / (sum(loads) / world) < 1.01           # LPT balance
    assert r["ids"] == lpt_partition(costs, world)            # every rank derives the same partition
