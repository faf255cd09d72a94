"""Worker for tests/test_multirank_cpu.py: run under torch.distributed.run with world size 2
on CPU (gloo).  Exercises the launcher plumbing bench.py uses for N > 1 and the ensemble
sharding / re-ordering logic; the NCCL gather itself is GPU-only and emulated with gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from enstop_b200 import enstop_  # noqa: E402


def main():
    rank, world, local = bench.dist_env()
    plumb = bench.Plumbing(rank, world)
    plumb.barrier()
    assert plumb.max(10.0 + rank) == 10.0 + world - 1
    assert plumb.sum(1.5) == 1.5 * world
    token = plumb.bcast_bytes(b"unique-id-from-rank0" if rank == 0 else None)
    assert token == b"unique-id-from-rank0"

    # ensemble sharding: 5 members over `world` ranks, k=3 topics, m=4 terms
    n_runs, k, m = 5, 3, 4
    shards = enstop_.shard_members(n_runs, world)
    seeds = enstop_.member_seeds(123, n_runs)          # same on every rank
    mine = shards[rank]
    local_topics = np.concatenate(
        [np.full((k, m), float(seeds[r] % 1000) + r, dtype=np.float32) for r in mine]) \
        if mine else np.zeros((0, m), dtype=np.float32)
    import torch.distributed as dist
    box = [None] * world if rank == 0 else None
    dist.gather_object(local_topics, box, dst=0)        # stands in for plsa_comm_gather_topics
    if rank == 0:
        stacked = np.concatenate(box)                   # rank-major, as the NCCL gather delivers
        out = enstop_.stack_in_member_order(stacked, shards, k)
        for r in range(n_runs):
            assert np.all(out[r * k:(r + 1) * k] == float(seeds[r] % 1000) + r), r
        print("MULTIRANK_OK", world)
    plumb.barrier()
    plumb.close()


if __name__ == "__main__":
    main()
