"""Run-to-run reproducibility of one forward per mode (same engine, same input)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from _fixtures import htdemucs_config, init_weights, synth_mix, rel_l2
from demucs_b200.engine import Engine
cfg = htdemucs_config()
W = init_weights(cfg, 0, layer_scale=0.5)
mix = synth_mix(2, cfg.segment_length, 3).to("cuda:0")
for mode in ("fp32", "tf32", "tf32x3", "strict", "bf16"):
    eng = Engine(cfg, W, "cuda:0", mode=mode)
    t1, t2 = {}, {}
    a = eng.forward(mix, t1).clone()
    b = eng.forward(mix, t2).clone()
    first = next((k for k in t1 if not torch.equal(t1[k], t2[k])), None)
    print(mode, "rel-L2 between two runs:", rel_l2(a.cpu(), b.cpu()), "first differing tap:", first,
          None if first is None else rel_l2(t1[first].cpu(), t2[first].cpu()), flush=True)
