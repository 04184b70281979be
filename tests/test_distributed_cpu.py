"""CPU, gloo, world_size 2: the host-side data-parallel logic (sample sharding + the single allreduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fireflies_b200.parallel import allreduce_sum_, fold_allreduce, max_over_ranks, shard_samples


def test_shard_samples_partitions_exactly():
    for total in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, n = shard_samples(total, r, world)
                seen += list(range(first, first + n))
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        shard_samples(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, n_pts):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds the same per-sample gradient table (a stand-in for the splat backward), sums its own
        # shard, and the allreduce must reproduce the full-batch sum independent of the split
        g = torch.Generator().manual_seed(0)
        per_sample = torch.randn(total, n_pts, 2, generator=g, dtype=torch.float64)
        first, n = shard_samples(total, rank, world)
        part = per_sample[first:first + n].sum(0)
        allreduce_sum_(part)
        assert torch.allclose(part, per_sample.sum(0), rtol=1e-12, atol=1e-12)
        assert max_over_ranks(float(rank + 1), "cpu") == float(world)
        # the fused fold + exchange is a GPU kernel (peer memory or fold + NCCL): host tensors are refused, not summed on the CPU
        with pytest.raises(RuntimeError, match="no CPU path"):
            fold_allreduce(per_sample[first:first + n].float())
    finally:
        dist.destroy_process_group()


def test_allreduce_of_sharded_pattern_gradients_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 37, 16), nprocs=2, join=True)
