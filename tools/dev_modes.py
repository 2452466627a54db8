"""Development check: parity of the tensor-core modes on the htdemucs golden fixtures + per-launch profile."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from _fixtures import golden, rel_l2, strided, forward_fixture_inputs, htdemucs_config
from demucs_b200.engine import Engine
from oracle.htdemucs_oracle import htdemucs_forward

modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["strict", "bf16"]
cfg = htdemucs_config()
for name in ("htdemucs_ls05.npz", "htdemucs_default.npz"):
    g = golden(name)
    W, mix = forward_fixture_inputs(g, cfg)
    for mode in modes:
        try:
            eng = Engine(cfg, W, "cuda:0", mode=mode)
            taps = {}
            got = eng.forward(mix.to("cuda:0"), taps)
            torch.cuda.synchronize()
        except Exception as e:
            print(name, mode, "FAILED", repr(e)[:300])
            continue
        errs = {k[4:]: rel_l2(strided(taps[k[4:]].contiguous(), int(g["tap_stride"])), g[k]) for k in g.files if k.startswith("tap.")}
        e_out = rel_l2(strided(got, int(g["stride"])), g["out"])
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"{name} {mode}: out {e_out:.3e}; worst taps {[(k, f'{v:.2e}') for k, v in worst]}", flush=True)
