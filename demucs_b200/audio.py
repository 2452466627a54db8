"""Audio front and back door on the device -- drop-in for the tensor side of reference ``demucs/audio.py``.

``convert_audio_channels`` / ``convert_audio`` (audio.py:143-172), ``prevent_clip`` (audio.py:218-233), ``i16_pcm``
(audio.py:175-180) and ``save_audio`` (audio.py:236-265) with the reference's names, arguments and error behaviour, for
CUDA tensors: channel conversion and ``julius.resample_frac`` as one polyphase kernel, clip prevention + PCM
quantisation + interleaving as one kernel, so that a stem crosses PCIe as 16-bit frames.  There is no CPU path:
CPU tensors raise ``KernelError`` (use the reference, or move the tensor to the GPU).

``julius`` is a third-party dependency that is not vendored in the reference tree (requirements.txt: julius>=0.2.3)
and not installed here; the filter bank below restates its published algorithm (julius/resample.py, ResampleFrac:
zeros=24, rolloff=0.945, Hann-windowed sinc, every phase normalised to unit sum, replicate padding, floor output
length) -- parity with julius itself is UNPINNED, see oracle/audio_oracle.py.  File decoding (ffmpeg / torchaudio)
and mp3 / flac encoding stay outside; ``save_audio`` writes RIFF/WAVE itself.
"""
from __future__ import annotations

import math
import struct
import typing as tp
from pathlib import Path

import torch

from . import _lib
from ._lib import ptr

_CLIP = {None: 0, "none": 0, "rescale": 1, "clamp": 2, "tanh": 3}
_KERNELS: tp.Dict[tp.Tuple, tp.Tuple[torch.Tensor, int, int, int]] = {}


def _require_cuda(wav: torch.Tensor, what: str) -> None:
    if wav.device.type != "cuda" and _lib.TEST_HOOK is None:
        raise _lib.KernelError(f"demucs_b200.audio.{what} runs on CUDA tensors only (there is no CPU path)")


def _stream(wav: torch.Tensor) -> int:
    return torch.cuda.current_stream(wav.device).cuda_stream if wav.device.type == "cuda" else 0


def resample_kernel(old_sr: int, new_sr: int, zeros: int = 24, rolloff: float = 0.945):
    """julius.ResampleFrac._init_kernels: ([new_sr, 2*width + old_sr] float32 filter bank, old_sr, new_sr, width), the
    rates reduced by their gcd."""
    gcd = math.gcd(old_sr, new_sr)
    old_sr, new_sr = old_sr // gcd, new_sr // gcd
    sr = min(new_sr, old_sr) * rolloff
    width = math.ceil(zeros * old_sr / sr)
    idx = torch.arange(-width, width + old_sr).float()
    kernels = []
    for i in range(new_sr):
        t = (-i / new_sr + idx / old_sr) * sr
        t = t.clamp_(-zeros, zeros)
        t *= math.pi
        window = torch.cos(t / zeros / 2) ** 2
        kernel = torch.where(t == 0, torch.ones_like(t), torch.sin(t) / t) * window
        kernel.div_(kernel.sum())
        kernels.append(kernel)
    return torch.stack(kernels).contiguous(), old_sr, new_sr, width


def convert_audio_channels(wav: torch.Tensor, channels: int = 2) -> torch.Tensor:
    """audio.py:143-166 for a CUDA tensor [..., src_channels, length]."""
    *shape, src_channels, length = wav.shape
    if src_channels == channels:
        return wav
    if not (channels == 1 or src_channels == 1 or src_channels >= channels):
        raise ValueError('The audio file has less channels than requested but is not mono.')
    _require_cuda(wav, "convert_audio_channels")
    x = wav.contiguous().float()
    items = int(math.prod(shape)) if shape else 1
    y = torch.empty(*shape, channels, length, dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device) if wav.device.type == "cuda" else _null():
        _lib.call("bd_convert_channels", ptr(x), ptr(y), items, src_channels, channels, length, _stream(wav))
    return y


def convert_audio(wav: torch.Tensor, from_samplerate: int, to_samplerate: int, channels: int) -> torch.Tensor:
    """audio.py:169-172: to ``channels`` channels, then ``julius.resample_frac(wav, from_samplerate, to_samplerate)``
    (output length floor(length * to / from)) -- one kernel for both."""
    *shape, src_channels, length = wav.shape
    if from_samplerate == to_samplerate:
        return convert_audio_channels(wav, channels)
    if not (src_channels == channels or channels == 1 or src_channels == 1 or src_channels >= channels):
        raise ValueError('The audio file has less channels than requested but is not mono.')
    _require_cuda(wav, "convert_audio")
    key = (int(from_samplerate), int(to_samplerate), str(wav.device))
    if key not in _KERNELS:
        k, o, n, w = resample_kernel(int(from_samplerate), int(to_samplerate))
        _KERNELS[key] = (k.to(wav.device), o, n, w)
    kern, old_sr, new_sr, width = _KERNELS[key]
    out_len = int(new_sr * length // old_sr)
    x = wav.contiguous().float()
    items = int(math.prod(shape)) if shape else 1
    y = torch.empty(*shape, channels, out_len, dtype=torch.float32, device=wav.device)
    if out_len:
        with torch.cuda.device(wav.device) if wav.device.type == "cuda" else _null():
            _lib.call("bd_resample_frac", ptr(x), ptr(y), ptr(kern), items, src_channels, channels, length, out_len, old_sr,
                      new_sr, width, _stream(wav))
    return y


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _clip_pcm(wav: torch.Tensor, clip, bits: int) -> torch.Tensor:
    """wav [channels, frames] (CUDA float) -> [frames, channels] int16 / int32(24-bit) / float32, clip prevention applied."""
    if clip not in _CLIP:
        raise ValueError(f"Invalid mode {clip}")
    if not wav.dtype.is_floating_point:
        raise AssertionError("too late for clipping")
    _require_cuda(wav, "prevent_clip / PCM")
    if wav.dim() != 2:
        raise ValueError("expected wav of shape [channels, frames]")
    x = wav.contiguous().float()
    C_, T = x.shape
    out = torch.empty(T, C_, dtype={16: torch.int16, 24: torch.int32, 32: torch.float32}[bits], device=wav.device)
    peak = torch.empty(1, dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device) if wav.device.type == "cuda" else _null():
        if _CLIP[clip] == 1:
            _lib.call("bd_absmax", ptr(x), ptr(peak), x.numel(), _stream(wav))
        if T:
            _lib.call("bd_clip_pcm", ptr(x), ptr(out), C_, T, _CLIP[clip], ptr(peak), bits, _stream(wav))
    return out


def prevent_clip(wav: torch.Tensor, mode: tp.Optional[str] = 'rescale') -> torch.Tensor:
    """audio.py:218-233 (``wav`` [channels, frames] or [..., frames]; returns the same layout)."""
    if mode is None or mode == 'none':
        return wav
    assert wav.dtype.is_floating_point, "too late for clipping"
    if mode not in _CLIP:
        raise ValueError(f"Invalid mode {mode}")
    flat = wav.reshape(1, -1)
    return _clip_pcm(flat, mode, 32).reshape(-1).view(wav.shape)      # one channel: interleaving is the identity


def i16_pcm(wav: torch.Tensor) -> torch.Tensor:
    """audio.py:175-180: float -> int16 (clamp to [-1, 1], * 32767, truncate); integer input is returned as is."""
    if not wav.dtype.is_floating_point:
        return wav
    return _clip_pcm(wav.reshape(1, -1), "none", 16).reshape(-1).view(wav.shape)


def stems_to_pcm(wav: torch.Tensor, clip: tp.Optional[str] = 'rescale', bits_per_sample: int = 16,
                 as_float: bool = False) -> torch.Tensor:
    """The back door: wav [channels, frames] on the GPU -> interleaved frames [frames, channels] on the HOST (pinned),
    clip prevention and quantisation done on the device, so 2 (or 4) bytes per sample cross PCIe."""
    bits = 32 if as_float else int(bits_per_sample)
    if bits == 32 and not as_float:
        raise ValueError("32-bit integer PCM is not produced; use as_float=True for 32-bit float")
    dev = _clip_pcm(wav, clip, bits)
    host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=dev.device.type == "cuda")
    host.copy_(dev, non_blocking=True)
    if dev.device.type == "cuda":
        torch.cuda.current_stream(dev.device).synchronize()
    return host


def save_audio(wav: torch.Tensor, path: tp.Union[str, Path], samplerate: int, bitrate: int = 320,
               clip: tp.Optional[str] = 'rescale', bits_per_sample: int = 16, as_float: bool = False, preset: int = 2):
    """audio.py:236-265 for ``.wav`` targets: clip prevention, quantisation and interleaving on the device, then a
    RIFF/WAVE file written directly (PCM_S 16 / 24 bit or IEEE float 32).  mp3 / flac need lameenc / torchaudio, which
    are outside this build."""
    path = Path(path)
    suffix = path.suffix.lower()
    if suffix in (".mp3", ".flac"):
        raise ValueError(f"{suffix} encoding is outside this build (lameenc / torchaudio); write .wav")
    if suffix != ".wav":
        raise ValueError(f"Invalid suffix for path: {suffix}")
    frames = stems_to_pcm(wav, clip, bits_per_sample, as_float)
    bits = 32 if as_float else int(bits_per_sample)
    n, ch = frames.shape
    if bits == 24:      # int32 container -> packed little-endian 3-byte samples
        data = frames.numpy().astype("<i4").view("u1").reshape(n, ch, 4)[..., :3].tobytes()
    else:
        data = frames.numpy().tobytes()
    block = ch * bits // 8
    fmt = 3 if as_float else 1
    header = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, fmt, ch, samplerate, samplerate * block, block, bits) + b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(header)
        f.write(data)
