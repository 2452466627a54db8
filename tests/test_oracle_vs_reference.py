"""CPU, build container only: the oracle against the live, unmodified reference
(skipped where /root/reference does not exist, e.g. on the GPU box)."""
import random

import pytest
import torch

from oracle import refload
from oracle.htdemucs_oracle import htdemucs_forward
from oracle.apply_oracle import apply_model_oracle
from demucs_b200.config import HTDemucsConfig
from demucs_b200.weights import param_specs, count_params
from _fixtures import rel_l2, small_config, synth_mix, init_weights, htdemucs_config

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not present")


def test_param_inventory_is_the_reference_state_dict():
    cfg = htdemucs_config()
    assert count_params(cfg) == 41984456
    ref = refload.load()
    torch.manual_seed(0)
    model = ref.HTDemucs(**small_config().reference_kwargs())
    sd = model.state_dict()
    specs = param_specs(small_config())
    assert list(sd.keys()) == list(specs.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(specs[k][0]), k
    assert HTDemucsConfig.from_reference_kwargs(*model._init_args_kwargs[0],
                                                **model._init_args_kwargs[1]) == small_config()


def test_forward_with_the_references_own_init():
    """Weights straight out of the reference constructor (torch.manual_seed(0))."""
    ref = refload.load()
    cfg = small_config()
    torch.manual_seed(0)
    model = ref.HTDemucs(**cfg.reference_kwargs()).eval()
    mix = synth_mix(2, cfg.segment_length - 777, 5)
    with torch.no_grad():
        want = model(mix)
        got = htdemucs_forward(model.state_dict(), cfg, mix)
    assert rel_l2(got, want) < 5e-6


def test_apply_model_shifts_and_rng_stream():
    ref = refload.load()
    cfg = small_config()
    W = init_weights(cfg, 3, layer_scale=0.5)
    model = refload.build_reference_model(cfg, W)
    mix = synth_mix(1, 44100 * 2 + 123, 17)
    with torch.no_grad():
        random.seed(42)
        want = ref.apply_model(model, mix.clone(), shifts=3, overlap=0.3)
        state_ref = random.getstate()
        random.seed(42)
        got = apply_model_oracle((W, cfg), mix, shifts=3, overlap=0.3)
    assert random.getstate() == state_ref  # same number of RNG draws
    assert rel_l2(got, want) < 5e-6


@pytest.mark.parametrize("length", [343980, 100001])
def test_hdemucs_oracle_matches_live_reference(length):
    """Hybrid Demucs v3 (hdemucs.py:689-794): inventory and forward of the oracle against the unmodified reference."""
    from demucs_b200 import hdemucs as HD
    from oracle.hdemucs_oracle import hdemucs_forward
    from oracle.make_golden import hdemucs_small_config, hdemucs_reference
    cfg = hdemucs_small_config()
    W = HD.init_weights(cfg, 3, 0.5)
    model = hdemucs_reference(cfg, W)
    assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == [(k, tuple(s[0])) for k, s in HD.param_specs(cfg).items()]
    mix = synth_mix(1, length, 9)
    with torch.no_grad():
        assert rel_l2(hdemucs_forward(W, cfg, mix), model(mix)) < 2e-6


def test_audio_oracle_matches_live_reference():
    import sys
    import types
    refload.load()
    sys.modules.setdefault("lameenc", types.ModuleType("lameenc"))
    import demucs.audio as ra
    from oracle import audio_oracle as O
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 777, generator=g) * 1.5
    for ch in (1, 2, 3):
        assert torch.equal(O.convert_audio_channels(x, ch), ra.convert_audio_channels(x, ch))
    for mode in ("rescale", "clamp", "tanh", "none"):
        assert torch.equal(O.prevent_clip(x.clone(), mode), ra.prevent_clip(x.clone(), mode))
    assert torch.equal(O.i16_pcm(x.clone()), ra.i16_pcm(x.clone()))
