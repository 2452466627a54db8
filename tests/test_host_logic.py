"""CPU: the product's host logic (weight packing, implicit-GEMM descriptors, buffer plumbing, the
apply_model batcher, BagOfModels, Separator) checked against the oracle / reference goldens with the
numpy ABI emulator standing in for the CUDA library (tests/abi_emulator.py)."""
import random

import pytest
import torch

from abi_emulator import emulated_abi, CALLS
from oracle.htdemucs_oracle import htdemucs_forward
from _fixtures import (golden, rel_l2, strided, small_config, synth_mix, init_weights, forward_fixture_inputs,
                       APPLY_CASES, BAG_WEIGHTS)
import demucs_b200 as D
from demucs_b200.engine import Engine


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_engine_forward_matches_oracle_and_golden(mode):
    """Both descriptor forms (the tf32 mode uses the 3-tap transposed-conv packing)."""
    g = golden("small_short.npz")
    cfg = small_config()
    W, mix = forward_fixture_inputs(g, cfg)
    taps_o, taps = {}, {}
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix, taps_o)
    with emulated_abi():
        eng = Engine(cfg, W, "cpu", mode=mode)
        got = eng.forward(mix, taps if mode == "fp32" else None)
    assert rel_l2(got, want) < 1e-5
    assert rel_l2(strided(got, int(g["stride"])), g["out"]) < 1e-5
    for k, v in taps.items():
        ref = taps_o[k][..., :v.shape[-1]] if k in ("istft", "time_out") else taps_o[k]
        assert rel_l2(v, ref) < 1e-5, k


def test_apply_model_and_bag_match_reference_golden():
    g = golden("apply_small.npz")
    cfg = small_config()
    models = [D.HTDemucs.from_config(cfg, init_seed=s, layer_scale=0.5) for s in range(2)]
    mix = synth_mix(1, int(g["length"]), 99)
    stride = int(g["stride"])
    with emulated_abi():
        for name in ("power2", "nosplit", "shifts2"):   # shifts + RNG stream + transition power; leaf branch; two shifts
            m = mix[..., :50000] if name == "nosplit" else mix
            random.seed(0)
            out = D.apply_model(models[0], m.clone(), **APPLY_CASES[name])
            assert rel_l2(strided(out, stride), g[name]) < 1e-5, name
        random.seed(0)
        out = D.apply_model(D.BagOfModels(models, BAG_WEIGHTS), mix.clone(), shifts=1)
        assert rel_l2(strided(out, stride), g["bag"]) < 1e-5


def test_callbacks_and_errors():
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0)
    mix = synth_mix(1, 90000, 1)
    events = []
    with emulated_abi():
        D.apply_model(model, mix, shifts=0, callback=events.append, callback_arg={"tag": 7})
        with pytest.raises(ValueError):
            model(torch.zeros(1, 2, cfg.segment_length + 1))
        with pytest.raises(AssertionError):
            D.apply_model(model, mix, transition_power=0.5)

        def boom(d):
            raise KeyboardInterrupt
        with pytest.raises(KeyboardInterrupt):
            D.apply_model(model, mix, shifts=0, callback=boom)
    # reference callback contract (api.py:108-115): start/end per segment, with the progress keys
    assert [e["state"] for e in events] == ["start", "start", "end", "end"]
    assert {e["segment_offset"] for e in events} == {0, int(0.75 * cfg.segment_length)}
    assert all(e["tag"] == 7 and e["models"] == 1 and e["shift_idx"] == 0 for e in events)
    with pytest.raises(NotImplementedError):
        D.BagOfModels([model]).forward(mix)
    with pytest.raises(D.UnsupportedConfig):
        D.HTDemucs(["a"], cac=False)
    with pytest.raises(D.KernelError):                    # no CPU / PyTorch fallback outside the emulator
        model(torch.zeros(1, 2, 4096))


def test_tensor_chunk_and_center_trim_follow_reference_semantics():
    x = torch.arange(20.).view(1, 1, 20)
    c = D.TensorChunk(x, 5, 6)
    assert c.shape == [1, 1, 6]
    assert c.padded(10)[0, 0].tolist() == list(range(3, 13))          # real neighbours, not zeros
    assert D.TensorChunk(x, 16, 10).padded(8)[0, 0].tolist() == [14., 15., 16., 17., 18., 19., 0., 0.]
    nested = D.TensorChunk(c, 2, 100)
    assert (nested.offset, nested.length) == (7, 4)
    assert D.center_trim(x, 15)[0, 0].tolist() == list(range(2, 17))  # odd surplus trimmed on the right
    with pytest.raises(ValueError):
        D.center_trim(x, 21)


def test_emulated_dconv_conv3_is_the_dilated_conv():
    """The numpy statement of bd_dconv_conv3 (full-size layers only, so the small-model tests never reach it)."""
    import numpy as np
    import abi_emulator as E
    B, T, Fr, C, hid, dil = 2, 11, 4, 48, 6, 2
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, Fr, C, generator=g)
    w = torch.randn(hid, C, 3, generator=g)
    b = torch.randn(hid, generator=g)
    w1 = w.permute(0, 2, 1).reshape(hid, 3 * C).contiguous()
    M = B * T * Fr
    h = torch.zeros(M, 8)
    sums = torch.zeros(B * Fr, 2, dtype=torch.float64)
    E.bd_dconv_conv3(x.data_ptr(), w1.data_ptr(), b.data_ptr(), h.data_ptr(), 8, sums.data_ptr(), M, C, hid, T * Fr, Fr, dil, 1, 0)
    want = torch.nn.functional.conv1d(x.permute(0, 2, 3, 1).reshape(B * Fr, C, T), w, b, padding=dil, dilation=dil)
    rows = want.reshape(B, Fr, hid, T).permute(0, 3, 1, 2).reshape(M, hid)
    assert rel_l2(h[:, :hid], rows) < 1e-5
    assert np.allclose(sums[:, 0].numpy(), want.reshape(B * Fr, -1).double().sum(1).numpy(), atol=1e-3)


@pytest.mark.parametrize("channel_major", [0, 1])
def test_emulated_encoder_conv0(channel_major):
    """The numpy statement of bd_encoder_conv0 (htdemucs first layers only; the small test model never reaches it)."""
    import abi_emulator as E
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    B, cout = 2, 48
    cin, I1, Jin = (2, 1, 101) if channel_major else (4, 3, 32)
    x = torch.randn(B, cin, Jin, generator=g) if channel_major else torch.randn(B, I1, Jin, cin, generator=g)
    xin = x if channel_major else x.permute(0, 1, 3, 2).reshape(B * I1, cin, Jin)
    Io = (Jin + 3) // 4
    w = torch.randn(cout, cin, 8, generator=g)
    b = torch.randn(cout, generator=g)
    norm = torch.zeros(B, 8)
    norm[:, 0], norm[:, 2] = 0.3, 1.7
    wp = w.permute(0, 2, 1).reshape(cout, 8 * cin).contiguous()
    out = torch.zeros(B, I1, Io, cout)
    E.bd_encoder_conv0(x.data_ptr(), channel_major, norm.data_ptr(), 8, wp.data_ptr(), b.data_ptr(), out.data_ptr(),
                       B, I1, Io, Jin, cin, cout, 1, 0)
    xn = F.pad((xin - 0.3) * 1.7, (2, 4 * Io + 6 - Jin))
    want = F.gelu(F.conv1d(xn, w, b, stride=4))[..., :Io].reshape(B, I1, cout, Io).permute(0, 1, 3, 2)
    assert rel_l2(out, want) < 1e-5


def test_engine_tf32_host_paths_with_48_channels():
    """The dedicated first-layer / DConv entry points are only taken with the htdemucs channel count (48 -> hid 6):
    run the engine's host logic in "tf32" mode on a short 48-channel model through the emulated ABI and compare with
    the oracle (the emulator computes in fp32, so this checks wiring, geometry and workspaces, not rounding)."""
    from fractions import Fraction
    from demucs_b200.config import HTDemucsConfig
    cfg = HTDemucsConfig(sources=["a", "b"], channels=48, dconv_mode=3, bottom_channels=128, t_heads=2, t_layers=1,
                         segment=Fraction(1, 2))
    cfg.validate()
    W = init_weights(cfg, 5, layer_scale=0.5)
    mix = synth_mix(2, cfg.segment_length, 12)
    with torch.no_grad():
        want = htdemucs_forward(W, cfg, mix)
    import abi_emulator as E
    with emulated_abi():
        E.CALLS.clear()
        eng = Engine(cfg, W, "cpu", mode="tf32")
        got = eng.forward(mix)
        calls = set(E.CALLS)
    assert {"bd_encoder_conv0", "bd_dconv_conv3", "bd_dconv_expand_stats", "bd_dconv_expand_update"} <= calls
    assert rel_l2(got, want) < 1e-5


def test_forward_core_matches_reference_golden():
    """HTDemucs.forward_core (htdemucs.py:662-759, the ONNX-export surface) through the engine's host logic."""
    g = golden("core_small.npz")
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5, mode="fp32")
    mix = synth_mix(2, cfg.segment_length, 1240)
    from oracle.htdemucs_oracle import stft_cac
    gen = torch.Generator().manual_seed(5)
    mag = stft_cac(mix, cfg.nfft)
    mag = mag + 0.05 * mag.std() * torch.randn(mag.shape, generator=gen)
    assert rel_l2(strided(mag, 211), g["mag"]) < 1e-5
    with emulated_abi():
        spec_out, time_out = model.forward_core(mag, mix)
    assert list(spec_out.shape) == list(g["spec_shape"]) and list(time_out.shape) == list(g["time_shape"])
    assert rel_l2(strided(spec_out, 211), g["spec_out"]) < 1e-5
    assert rel_l2(strided(time_out, 7), g["time_out"]) < 1e-5
    with emulated_abi(), pytest.raises(ValueError):
        model.forward_core(mag[..., :-1], mix)
