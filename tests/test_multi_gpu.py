"""N > 1 on real devices: the peer-memory fold + allreduce kernel (ffb_fold_allreduce), sharding-invariant randomisation and
the allreduced pattern gradient against the single-rank full batch.  Self-spawns one process per
GPU (NCCL, 127.0.0.1); skipped on a box with a single device -- there `bench.py --gpus N` runs the same check
(`multi_gpu_check` in its JSON line) whenever the driver launches it with N > 1."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, symm):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FFB_SYMM_ALLREDUCE=symm)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        _checks(rank, world, dev, symm)
    except BaseException:  # noqa: BLE001  (a failing rank must not wait for its peers in destroy_process_group: leave at once)
        import traceback
        traceback.print_exc()
        os._exit(1)
    dist.destroy_process_group()


def _checks(rank, world, dev, symm):
    if True:
        from fireflies_b200.parallel import multi_gpu_selfcheck
        res = multi_gpu_selfcheck(dev)
        assert res["allreduce"] == res["rng_split_invariant"] == res["sharded_gradient"] == "ok" and res["ranks"] == world
        if symm == "0":
            assert res["exchange"] == "nccl"
        # a rank without samples contributes zeros and the epochs stay in step
        from fireflies_b200.parallel import check_folders, fold_allreduce
        x = torch.full((3 if rank == 0 else 0, 64, 2), 1.5, device=dev)
        for _ in range(3):
            y = fold_allreduce(x)
            assert torch.equal(y, torch.full((64, 2), 4.5, device=dev))
        check_folders()
        torch.cuda.synchronize()


@pytest.mark.parametrize("symm", ["1", "0"])
def test_two_rank_selfcheck(symm):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (bench.py --gpus N carries the same check as `multi_gpu_check`)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), symm), nprocs=world, join=True)


def test_selfcheck_single_rank():
    """The same three checks without a process group (one rank): exercises the code the multi-rank runs share."""
    from fireflies_b200.parallel import multi_gpu_selfcheck
    res = multi_gpu_selfcheck(torch.device("cuda", 0))
    assert res == {"rng_split_invariant": "ok", "allreduce": "ok", "sharded_gradient": "ok", "ranks": 1, "exchange": "single rank"}
