"""CPU (ABI emulator): StreamSeparator -- segment-by-segment separation with carried overlap-add state -- hands out,
piece by piece, exactly what apply_model(shifts=0, split=True) computes on the whole track."""
import random

import pytest
import torch

import demucs_b200 as D
from demucs_b200.streaming import StreamSeparator
from _fixtures import small_config, synth_mix
from abi_emulator import emulated_abi


@pytest.mark.parametrize("overlap,length", [(0.25, 150000), (0.6, 110000), (0.25, 40000)])
def test_stream_equals_apply_model(overlap, length):
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5, mode="fp32")
    mix = synth_mix(1, length, 3)
    rng = random.Random(5)
    with emulated_abi():
        want = D.apply_model(model, mix.clone(), shifts=0, split=True, overlap=overlap)[0]
        sep = StreamSeparator(model, overlap=overlap, device="cpu")
        pieces, pos, first_at = [], 0, None
        while pos < length:
            n = min(length - pos, rng.randrange(1, 30000))
            got = sep.push(mix[0, :, pos:pos + n])
            pos += n
            if got.shape[-1] and first_at is None:
                first_at = pos
            pieces.append(got)
        pieces.append(sep.flush())
        # the separator is reusable after a flush
        again = torch.cat([sep.push(mix[0]), sep.flush()], dim=-1)
    out = torch.cat(pieces, dim=-1)
    assert out.shape == want.shape
    assert (out - want).abs().max() <= 1e-6 * want.abs().max()
    assert (again - want).abs().max() <= 1e-6 * want.abs().max()
    if length > cfg.segment_length:       # the first stems leave as soon as one segment is in, not at the end
        assert first_at is not None and first_at < cfg.segment_length + 30000


@pytest.mark.gpu
def test_stream_equals_apply_model_gpu():
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5, mode="fp32").to("cuda:0")
    mix = synth_mix(1, 190000, 3)
    want = D.apply_model(model, mix.to("cuda:0"), shifts=0, split=True, overlap=0.25)[0]
    sep = StreamSeparator(model, overlap=0.25)
    pieces, pos = [], 0
    for n in (50000, 1, 90000, 20000, 29999):
        pieces.append(sep.push(mix[0, :, pos:pos + n]))
        pos += n
    pieces.append(sep.flush())
    out = torch.cat(pieces, dim=-1)
    assert out.shape == want.shape and (out - want).abs().max() <= 2e-5 * want.abs().max()
