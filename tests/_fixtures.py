"""Shared helpers for the parity tests (fixtures are defined by oracle/make_golden.py)."""
import os

import numpy as np
import torch

from oracle.make_golden import GOLDEN_DIR, small_config, synth_mix, APPLY_CASES, BAG_WEIGHTS  # noqa
from demucs_b200.config import htdemucs_config  # noqa
from demucs_b200.weights import init_weights  # noqa


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def strided(x, stride):
    return x.detach().reshape(-1)[::stride].float().cpu().numpy()


def forward_fixture_inputs(g, cfg):
    """(weights, mix) that produced golden file ``g``."""
    ls = float(g["layer_scale"])
    W = init_weights(cfg, int(g["seed"]), layer_scale=None if ls < 0 else ls)
    mix = synth_mix(int(g["batch"]), int(g["length"]), 1234 + int(g["seed"]))
    return W, mix
