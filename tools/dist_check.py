"""torchrun --nproc-per-node N tools/dist_check.py: sharded apply_model (NCCL) against the same call unsharded on
every rank, for the three exchange paths: heads only (shifts=0), slivers (shifts=2, 6 stems), bag members."""
import os, sys, random, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import demucs_b200 as D
from demucs_b200.distributed import Shard
from demucs_b200.config import htdemucs_6s_config, htdemucs_config

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "strict"
g = torch.Generator().manual_seed(5)
res = {}


def check(name, model, mix, **kw):
    random.seed(3)
    want = D.apply_model(model, mix.to(dev), device=dev, **kw)
    for gather in ("all", "none", "root"):
        random.seed(3 + 17 * rank)      # different draws per rank: rank 0's must win
        sh = Shard(gather=gather)
        got = D.apply_model(model, mix.to(dev), device=dev, shard=sh, **kw)
        a, b = sh.owned
        err = float((got[..., a:b] - want[..., a:b]).abs().max() / want.abs().max()) if b > a else 0.0
        t = torch.tensor([err], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"{name}/{gather}"] = float(t)
    # host path: pinned input, each rank gets its own range back
    sh = Shard(gather="none")
    random.seed(3)
    got = D.apply_model(model, mix.pin_memory(), device=dev, shard=sh, **kw)
    a, b = sh.owned
    err = float((got[..., a:b] - want[..., a:b].cpu()).abs().max() / want.abs().max()) if b > a else 0.0
    t = torch.tensor([err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[f"{name}/host"] = float(t)


m4 = D.htdemucs(mode=mode).to(dev)
check("shifts0_21seg", m4, 0.1 * torch.randn(1, 2, 21 * 257985, generator=g), shifts=0, overlap=0.25, batch_size=4)
check("shifts0_3seg", m4, 0.1 * torch.randn(1, 2, 3 * 257985 - 1000, generator=g), shifts=0, overlap=0.25)
m6 = D.HTDemucs.from_config(htdemucs_6s_config(), init_seed=2, mode=mode).to(dev)
check("6s_shifts2", m6, 0.1 * torch.randn(1, 2, 19 * 257985 + 777, generator=g), shifts=2, overlap=0.25, batch_size=8)
bag = D.BagOfModels([D.HTDemucs.from_config(htdemucs_config(), init_seed=10 + m, mode=mode).to(dev) for m in range(2)],
                    [[1., 0., 1., 0.5], [0., 1., 1., 0.5]])
check("bag2_shifts1", bag, 0.1 * torch.randn(1, 2, 9 * 257985, generator=g), shifts=1, overlap=0.25, batch_size=8)
if rank == 0:
    print(json.dumps({"world": world, "mode": mode, "max_rel_abs_err": res}))
dist.destroy_process_group()
