"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by ``demucs_b200``.

CPU restatement of the tensor side of reference demucs/audio.py for the separator's front / back door:
``convert_audio_channels`` (audio.py:143-166), ``prevent_clip`` (:218-233), ``i16_pcm`` (:175-180) -- pinned against the
reference's own functions by tests/golden/audio.npz (oracle/make_golden.py: audio_fixture) -- and ``resample_frac``.

``convert_audio`` (audio.py:169-172) delegates the resampling to ``julius.resample_frac``.  julius is a third-party
dependency (requirements.txt: julius>=0.2.3), absent from /root/reference and not installed in this image, so the
resampler below restates its published algorithm (julius/resample.py ``ResampleFrac``: zeros=24, rolloff=0.945, Hann-
windowed sinc evaluated at (idx/old_sr - i/new_sr)*min(old,new)*rolloff, clamped to +-zeros, each phase normalised to
unit sum; input replicate-padded by (width, width + old_sr); strided conv1d; output trimmed to floor(new*L/old)).
PARITY UNPINNED for this one function: there is neither a julius golden vector nor julius itself to run here; the
reference's only call site is audio.py:172 / api.py:266.  Its properties (DC and in-band sinusoids preserved, identity
for equal rates) are what the tests check.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def convert_audio_channels(wav: torch.Tensor, channels: int = 2) -> torch.Tensor:
    *shape, src, length = wav.shape
    if src == channels:
        return wav
    if channels == 1:
        return wav.mean(dim=-2, keepdim=True)
    if src == 1:
        return wav.expand(*shape, channels, length)
    if src >= channels:
        return wav[..., :channels, :]
    raise ValueError('The audio file has less channels than requested but is not mono.')


def resample_frac(x: torch.Tensor, old_sr: int, new_sr: int, zeros: int = 24, rolloff: float = 0.945) -> torch.Tensor:
    gcd = math.gcd(old_sr, new_sr)
    old_sr, new_sr = old_sr // gcd, new_sr // gcd
    if old_sr == new_sr:
        return x
    sr = min(new_sr, old_sr) * rolloff
    width = math.ceil(zeros * old_sr / sr)
    idx = torch.arange(-width, width + old_sr).float()           # julius builds its filters in float32
    kernels = []
    for i in range(new_sr):
        t = (-i / new_sr + idx / old_sr) * sr
        t = t.clamp_(-zeros, zeros)
        t *= math.pi
        window = torch.cos(t / zeros / 2) ** 2
        k = torch.where(t == 0, torch.ones_like(t), torch.sin(t) / t) * window
        k.div_(k.sum())
        kernels.append(k)
    kernel = torch.stack(kernels).view(new_sr, 1, -1).to(x.dtype)
    shape, length = x.shape, x.shape[-1]
    y = F.pad(x.reshape(-1, length)[:, None], (width, width + old_sr), mode='replicate')
    y = F.conv1d(y, kernel, stride=old_sr).transpose(1, 2).reshape(list(shape[:-1]) + [-1])
    return y[..., :int(new_sr * length // old_sr)]


def convert_audio(wav, from_samplerate, to_samplerate, channels):
    return resample_frac(convert_audio_channels(wav, channels), from_samplerate, to_samplerate)


def prevent_clip(wav: torch.Tensor, mode='rescale') -> torch.Tensor:
    if mode is None or mode == 'none':
        return wav
    assert wav.dtype.is_floating_point, "too late for clipping"
    if mode == 'rescale':
        return wav / max(1.01 * wav.abs().max(), 1)
    if mode == 'clamp':
        return wav.clamp(-0.99, 0.99)
    if mode == 'tanh':
        return torch.tanh(wav)
    raise ValueError(f"Invalid mode {mode}")


def i16_pcm(wav: torch.Tensor) -> torch.Tensor:
    if wav.dtype.is_floating_point:
        return (wav.clamp(-1, 1) * (2 ** 15 - 1)).short()
    return wav
