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


def _worker(rank, world, port, kwargs, length, ret, gather="all"):
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
        random.seed(7 + 1000 * rank)       # ranks draw DIFFERENT shift offsets: rank 0's must win on all of them
        shard = Shard(gather=gather)
        out = D.apply_model(model, mix, shard=shard, **kwargs)
    if gather == "none":                    # every rank reports the range it owns
        a, b = shard.owned
        ret.put((rank, a, b, out[..., a:b].numpy()))
    elif rank == 0:
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
    (dict(shifts=0, overlap=0.25), 160000),       # 4 segments -> 2 + 2, one halo segment
    (dict(shifts=1, overlap=0.6), 100000),        # heavy overlap: 2 halo segments, shifted window, RNG in step
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


@pytest.mark.parametrize("gather", ["none", "root"])
def test_sharded_apply_gather_policies(gather):
    """gather="none": the ranks' owned ranges tile the track and hold the single-process values (shifts=2: the
    ranges of the two passes differ, the slivers travel to their owners); gather="root": rank 0 has everything."""
    kwargs, length = dict(shifts=2, overlap=0.25), 110000
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kwargs, length, ret, gather)) for r in range(2)]
    for p in procs:
        p.start()
    got = [ret.get() for _ in range(2 if gather == "none" else 1)]
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    want = _single(kwargs, length)
    if gather == "root":
        assert abs(got[0] - want).max() <= 1e-6 * abs(want).max()
        return
    got.sort(key=lambda t: t[0])
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == length
    for _, a, b, piece in got:
        assert abs(piece - want[..., a:b]).max() <= 1e-6 * abs(want).max()


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


def test_sliver_schedule_routes_every_contribution_to_its_owner():
    """Pure arithmetic of Shard.combine on the real plans of multi-pass runs (shift trick on 10-minute and short tracks,
    1..8 ranks): after the scheduled pieces are added, every owner holds exactly one contribution per pass for every
    sample of its range, ownership tiles the track, and nothing is sent twice."""
    import numpy as np
    from demucs_b200.distributed import Shard

    class Fake(Shard):
        def __init__(self, rank, world):
            self.rank, self.world, self.group = rank, world, None
    rng = random.Random(3)
    seg_len, max_shift = 343980, 22050
    for world in (1, 2, 3, 8):
        for L in (26460000 // 40, 3 * 257985 - 1000, 500000, 7 * 257985 + 12345):
            for overlap in (0.25, 0.6):
                stride = int((1 - overlap) * seg_len)
                passes = []
                for _ in range(3):
                    offset = rng.randint(0, max_shift)
                    length = L + max_shift - offset
                    passes.append((-(-length // stride), seg_len, stride, length, max_shift - offset))
                sh = Fake(0, world)
                own, sched = sh.sliver_schedule(passes, L)
                assert own[0][0] == 0 and own[-1][1] == L or any(b > a for a, b in own)
                cover = np.zeros(L, np.int32)
                for a, b in own:
                    cover[a:b] += 1
                assert (cover == 1).all()
                have = [np.zeros(L, np.int32) for _ in range(world)]
                for q in range(world):
                    for p in passes:
                        a, b = sh.pass_range(q, p, L)
                        have[q][a:b] += 1
                assert len(set(sched)) == len(sched)
                got = [h.copy() for h in have]
                for q, r, x, y in sched:
                    assert own[r][0] <= x < y <= own[r][1] and not (own[q][0] <= x < own[q][1])
                    got[r][x:y] += have[q][x:y]
                for q, (a, b) in enumerate(own):
                    assert (got[q][a:b] == len(passes)).all(), (world, L, overlap, q)
