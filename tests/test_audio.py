"""Front / back door of the separator (demucs_b200/audio.py, csrc/audio.cu): the oracle against the reference's own
functions (tests/golden/audio.npz), the host logic through the ABI emulator (CPU), the kernels against the oracle (GPU)."""
import math
import struct

import numpy as np
import pytest
import torch

from _fixtures import golden, rel_l2
from abi_emulator import emulated_abi
from demucs_b200 import audio as A
from oracle import audio_oracle as O


def _inputs():
    g = torch.Generator().manual_seed(21)
    x5 = torch.randn(3, 5, 1000, generator=g)
    loud = 1.7 * torch.randn(2, 5000, generator=g)
    quiet = 0.2 * torch.randn(2, 5000, generator=g)
    return x5, loud, quiet


def test_oracle_matches_reference_golden():
    G = golden("audio.npz")
    x5, loud, quiet = _inputs()
    assert np.array_equal(O.convert_audio_channels(x5, 2).numpy(), G["ch_5to2"])
    assert np.allclose(O.convert_audio_channels(x5, 1).numpy(), G["ch_5to1"], atol=1e-7)
    assert np.array_equal(O.convert_audio_channels(x5[:, :1], 2).numpy(), G["ch_1to2"])
    with pytest.raises(ValueError):
        O.convert_audio_channels(x5[:, :2], 3)
    for mode in ("rescale", "clamp", "tanh"):
        assert np.allclose(O.prevent_clip(loud.clone(), mode).numpy(), G[f"clip_{mode}_loud"], atol=1e-7)
        assert np.allclose(O.prevent_clip(quiet.clone(), mode).numpy(), G[f"clip_{mode}_quiet"], atol=1e-7)
    assert np.array_equal(O.i16_pcm(loud.clone()).numpy(), G["i16_loud"])


def test_resampler_oracle_properties():
    """julius is absent (parity unpinned, oracle/audio_oracle.py): the restated algorithm must at least be a resampler --
    unit DC gain, an in-band sinusoid comes out at the new rate with its amplitude, equal rates are the identity."""
    sr_in, sr_out, n = 48000, 44100, 48000
    t = torch.arange(n, dtype=torch.float64) / sr_in
    x = (0.5 + 0.25 * torch.sin(2 * math.pi * 1000.0 * t))[None].float()
    y = O.resample_frac(x, sr_in, sr_out)
    assert y.shape[-1] == 44100
    t2 = torch.arange(y.shape[-1], dtype=torch.float64) / sr_out
    want = (0.5 + 0.25 * torch.sin(2 * math.pi * 1000.0 * t2)).float()
    assert (y[0, 200:-200] - want[200:-200]).abs().max() < 2e-4
    assert O.resample_frac(x, 44100, 44100) is x
    up = O.resample_frac(x, 22050, 44100)
    assert up.shape[-1] == 2 * n


def _cases():
    return [(48000, 44100, 2, 2, 30001), (22050, 44100, 1, 2, 9999), (44100, 16000, 5, 1, 20000), (32000, 44100, 3, 2, 4097)]


def _check_front_door(dev):
    for sr_in, sr_out, src, dst, n in _cases():
        g = torch.Generator().manual_seed(n)
        x = torch.randn(2, src, n, generator=g)
        got = A.convert_audio(x.to(dev), sr_in, sr_out, dst).cpu()
        want = O.convert_audio(x.double(), sr_in, sr_out, dst).float()
        assert got.shape == want.shape, (got.shape, want.shape)
        assert rel_l2(got, want) < 2e-6, (sr_in, sr_out, src, dst)
    x = torch.randn(3, 5, 1000)
    assert torch.equal(A.convert_audio_channels(x.to(dev), 2).cpu(), x[:, :2])
    assert torch.allclose(A.convert_audio_channels(x.to(dev), 1).cpu(), x.mean(1, keepdim=True), atol=1e-6)
    assert torch.equal(A.convert_audio(x[:, :1].to(dev), 44100, 44100, 2).cpu(), x[:, :1].expand(3, 2, 1000))
    with pytest.raises(ValueError):
        A.convert_audio_channels(x[:, :2].to(dev), 3)


def _check_back_door(dev, tmp_path):
    G = golden("audio.npz")
    _, loud, quiet = _inputs()
    for mode in ("rescale", "clamp", "tanh"):
        for name, x in (("loud", loud), ("quiet", quiet)):
            got = A.prevent_clip(x.to(dev), mode).cpu().numpy()
            assert np.allclose(got, G[f"clip_{mode}_{name}"], atol=2e-7, rtol=1e-6), (mode, name)
    assert A.prevent_clip(loud, "none") is loud
    with pytest.raises(ValueError):
        A.prevent_clip(loud.to(dev), "loudest")
    assert np.array_equal(A.i16_pcm(loud.to(dev)).cpu().numpy(), G["i16_loud"])
    # the wire format: clip prevention + quantisation + interleaving in one kernel
    frames = A.stems_to_pcm(loud.to(dev), "rescale", 16)
    want = O.i16_pcm(O.prevent_clip(loud.clone(), "rescale")).t()
    assert frames.dtype == torch.int16 and list(frames.shape) == [5000, 2]
    assert (frames.int() - want.int()).abs().max() <= 1          # a 1-ulp difference of the scale can flip a truncation
    f24 = A.stems_to_pcm(quiet.to(dev), "clamp", 24)
    assert f24.dtype == torch.int32 and (f24.float() / 8388607.0 - quiet.clamp(-0.99, 0.99).t()).abs().max() < 2e-7
    # .wav writer: header fields and payload
    path = tmp_path / "stem.wav"
    A.save_audio(loud.to(dev), path, 44100, clip="rescale", bits_per_sample=16)
    raw = path.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    fmt, ch, sr, _, block, bits = struct.unpack("<HHIIHH", raw[20:36])
    assert (fmt, ch, sr, block, bits) == (1, 2, 44100, 4, 16)
    assert struct.unpack("<I", raw[40:44])[0] == 5000 * 4 and len(raw) == 44 + 5000 * 4
    assert np.array_equal(np.frombuffer(raw[44:], dtype="<i2").reshape(5000, 2), frames.numpy())
    A.save_audio(quiet.to(dev), tmp_path / "f.wav", 44100, clip="none", as_float=True)
    raw = (tmp_path / "f.wav").read_bytes()
    assert struct.unpack("<H", raw[20:22])[0] == 3 and np.allclose(np.frombuffer(raw[44:], dtype="<f4").reshape(5000, 2), quiet.t())
    A.save_audio(quiet.to(dev), tmp_path / "p24.wav", 44100, clip="none", bits_per_sample=24)
    assert len((tmp_path / "p24.wav").read_bytes()) == 44 + 5000 * 2 * 3
    with pytest.raises(ValueError):
        A.save_audio(loud.to(dev), tmp_path / "x.mp3", 44100)


def test_host_logic_front_and_back_door(tmp_path):
    with emulated_abi():
        _check_front_door("cpu")
        _check_back_door("cpu", tmp_path)


def test_cpu_tensors_are_refused_without_a_gpu_path():
    from demucs_b200 import _lib
    with pytest.raises(_lib.KernelError):
        A.convert_audio(torch.zeros(1, 2, 100), 48000, 44100, 2)


@pytest.mark.gpu
def test_gpu_front_and_back_door(tmp_path):
    _check_front_door("cuda:0")
    _check_back_door("cuda:0", tmp_path)


@pytest.mark.gpu
def test_gpu_separator_resamples_and_emits_pcm():
    """Separator on 48 kHz mono input: converted on the device (api.py:265-266), separated, and taken off the GPU as
    int16 frames; compared with the oracle's convert -> normalise -> apply -> de-normalise -> clip -> quantise chain."""
    import demucs_b200 as D
    from _fixtures import small_config, init_weights
    from oracle.apply_oracle import apply_model_oracle
    cfg = small_config()
    model = D.HTDemucs.from_config(cfg, init_seed=0, layer_scale=0.5)
    sep = D.Separator(model, device="cuda:0", shifts=0)
    g = torch.Generator().manual_seed(8)
    wav48 = 0.3 * torch.randn(1, 90000, generator=g)
    _, stems = sep.separate_tensor(wav48.clone(), sr=48000)
    wav = O.convert_audio(wav48, 48000, 44100, 2)
    ref = wav.mean(0)
    x = (wav - ref.mean()) / (ref.std() + 1e-8)
    with torch.no_grad():
        want = apply_model_oracle((init_weights(cfg, 0, layer_scale=0.5), cfg), x[None], shifts=0)[0]
    want = want * (ref.std() + 1e-8) + ref.mean()
    for i, s in enumerate(cfg.sources):
        assert stems[s].shape == want[i].shape and rel_l2(stems[s], want[i]) < 1e-4
    pcm = sep.separate_tensor_pcm(wav48.clone(), sr=48000, clip="rescale", bits_per_sample=16)
    for i, s in enumerate(cfg.sources):
        w16 = O.i16_pcm(O.prevent_clip(want[i].clone(), "rescale")).t()
        assert pcm[s].dtype == torch.int16 and pcm[s].shape == w16.shape
        assert (pcm[s].int() - w16.int()).abs().float().mean() < 0.6       # quantisation steps, not signal differences


def test_separator_front_and_back_door_host_logic():
    """Separator on a v3 model object with 48 kHz mono input, through the emulated ABI: converted (api.py:265-266),
    separated, returned as float stems and as int16 frames; against the oracle chain."""
    import demucs_b200 as D
    from demucs_b200 import hdemucs as HD
    from oracle.apply_oracle import apply_model_oracle
    from oracle.make_golden import hdemucs_small_config
    cfg = hdemucs_small_config()
    cfg.segment = 2.0
    model = HD.HDemucs.from_config(cfg, init_seed=1, layer_scale=0.5, mode="fp32")
    g = torch.Generator().manual_seed(8)
    wav48 = 0.3 * torch.randn(1, 60000, generator=g)
    with emulated_abi():
        sep = D.Separator(model, device="cpu", shifts=0)
        _, stems = sep.separate_tensor(wav48.clone(), sr=48000)
        pcm = sep.separate_tensor_pcm(wav48.clone(), sr=48000, clip="clamp", bits_per_sample=16)
    wav = O.convert_audio(wav48, 48000, 44100, 2)
    ref = wav.mean(0)
    x = (wav - ref.mean()) / (ref.std() + 1e-8)
    with torch.no_grad():
        want = apply_model_oracle((HD.init_weights(cfg, 1, 0.5), cfg), x[None], shifts=0)[0]
    want = want * (ref.std() + 1e-8) + ref.mean()
    for i, s in enumerate(cfg.sources):
        assert stems[s].shape == want[i].shape and rel_l2(stems[s], want[i]) < 2e-5
        w16 = O.i16_pcm(O.prevent_clip(want[i].clone(), "clamp")).t()
        assert pcm[s].dtype == torch.int16 and (pcm[s].int() - w16.int()).abs().max() <= 1
