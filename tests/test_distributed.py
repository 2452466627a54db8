"""CPU, world_size 2 over gloo: the segment sharder (demucs_b200/distributed.py) with the numpy ABI emulator
standing in for the kernels.  Sharded apply_model must reproduce the single-process result exactly: each rank
overlap-adds only the samples it owns, after receiving its left neighbour's halo segments."""
import os
import random
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, kwargs, length, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from abi_emulator import emulated_abi
    from _fixtures import small_config, synth_mix
    import demucs_b200 as D
    from demucs_b200.distributed import Shard
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5)
    mix = synth_mix(1, length, 3)
    with emulated_abi():
        random.seed(7)
        out = D.apply_model(model, mix, shard=Shard(), **kwargs)
    if rank == 0:
        ret.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _single(kwargs, length):
    from abi_emulator import emulated_abi
    from _fixtures import small_config, synth_mix
    import demucs_b200 as D
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5)
    mix = synth_mix(1, length, 3)
    with emulated_abi():
        random.seed(7)
        return D.apply_model(model, mix, **kwargs).numpy()


@pytest.mark.parametrize("kwargs,length", [
    (dict(shifts=0, overlap=0.25), 190000),       # 4 segments -> 2 + 2, one halo segment
    (dict(shifts=1, overlap=0.6), 120000),        # heavy overlap: 2 halo segments, shifted window, RNG in step
    (dict(shifts=0, overlap=0.25), 60000),        # one segment on two ranks: rank 1 owns nothing (empty gather piece)
])
def test_sharded_apply_matches_single_process(kwargs, length):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kwargs, length, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    want = _single(kwargs, length)
    assert got.shape == want.shape
    assert abs(got - want).max() <= 1e-6 * abs(want).max()


def test_block_partition_and_halo():
    from demucs_b200.distributed import Shard

    class Fake(Shard):
        def __init__(self, rank, world):
            self.rank, self.world, self.group = rank, world, None
    blocks = [Fake(r, 8).block(103) for r in range(8)]     # 13 x 7 + 12
    assert [h - l for l, h in blocks] == [13] * 7 + [12]
    assert blocks[0][0] == 0 and blocks[-1][1] == 103
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(7))
    assert Shard.halo(343980, 257985) == 1 and Shard.halo(100, 40) == 2 and Shard.halo(100, 100) == 0
    assert [Fake(r, 4).block(2) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
