"""Hybrid Demucs v3 (``HDemucs``, reference demucs/hdemucs.py:338-794) -- configuration, parameter inventory and the
model object for the ``hdemucs_mmi`` architecture family (BASELINE configs[1]): six-layer dual U-Net without a
transformer, GroupNorm(4) in the two innermost layers, BiLSTM + LocalState attention inside their DConv branches,
the time branch injected into the frequency branch at layer 4 and split off again in the decoder.

The layer table below restates the constructor loop (hdemucs.py:470-585); the arithmetic runs in
``demucs_b200.hdemucs_engine.HDemucsEngine`` (sm_100a kernels, no PyTorch fallback).  Outside the accelerated path
(rejected loudly): Wiener filtering (``cac=False``), ``hybrid=False`` / ``hybrid_old``, ``multi_freqs`` (MultiWrap),
``channels_time``.
"""
from __future__ import annotations

import collections
import math
import typing as tp
from dataclasses import dataclass, field

import torch
from torch import nn

from .config import UnsupportedConfig

_PINNED = {"cac": True, "hybrid": True, "hybrid_old": False, "multi_freqs": None, "wiener_iters": 0, "end_iters": 0,
           "wiener_residual": False, "rewrite": True, "channels_time": None, "growth": 2, "nfft": 4096, "kernel_size": 8,
           "stride": 4, "time_stride": 2, "context": 1, "context_enc": 0, "depth": 6}
_IGNORED = {"multi_freqs_depth", "emb_smooth", "rescale"}


@dataclass
class HDemucsConfig:
    sources: tp.List[str] = field(default_factory=lambda: ["drums", "bass", "other", "vocals"])
    audio_channels: int = 2
    channels: int = 48
    nfft: int = 4096
    depth: int = 6
    freq_emb: float = 0.2
    emb_scale: float = 10.0
    norm_starts: int = 4
    norm_groups: int = 4
    dconv_mode: int = 1
    dconv_depth: int = 2
    dconv_comp: int = 4
    dconv_attn: int = 4
    dconv_lstm: int = 4
    dconv_init: float = 1e-4
    samplerate: int = 44100
    segment: tp.Union[float, int] = 40

    @property
    def hop(self) -> int:
        return self.nfft // 4

    @property
    def n_sources(self) -> int:
        return len(self.sources)

    def frames(self, length: int) -> int:
        return int(math.ceil(length / self.hop))

    def layers(self) -> tp.List[dict]:
        """Per encoder index: the geometry decided by hdemucs.py:470-585 (decoder j = depth-1-index mirrors it)."""
        A, S = self.audio_channels, self.n_sources
        chin, chin_z = A, 2 * A
        chout = chout_z = self.channels
        freqs = self.nfft // 2
        out = []
        for index in range(self.depth):
            freq = freqs > 1
            ker, stri = (8, 4) if freq else (4, 2)
            pad, last_freq = True, False
            if freq and freqs <= 8:
                ker, pad, last_freq = freqs, False, True
            if last_freq:
                chout_z = max(chout, chout_z)
                chout = chout_z
            out.append(dict(index=index, freq=freq, ker=ker, stride=stri, pad=pad, last_freq=last_freq,
                            norm=index >= self.norm_starts, lstm=index >= self.dconv_lstm, attn=index >= self.dconv_attn,
                            chin=chin, chout=chout, chin_z=chin_z, chout_z=chout_z, freqs_in=freqs,
                            has_time=freq, time_empty=last_freq))
            if index == 0:
                chin = A * S
                chin_z = 2 * chin
            # the decoder of this index maps chout(_z) back to (this) chin(_z)
            out[-1]["dec_out"], out[-1]["dec_out_z"] = chin, chin_z
            chin, chin_z = chout, chout_z
            chout, chout_z = 2 * chout, 2 * chout_z
            if freq:
                freqs = 1 if freqs <= 8 else freqs // 4
        return out

    def validate(self) -> None:
        if self.nfft != 4096 or self.depth != 6:
            raise UnsupportedConfig("only nfft=4096, depth=6 is built (the hdemucs_mmi geometry)")
        if self.channels % (4 * self.dconv_comp) or self.channels % 8:
            raise UnsupportedConfig("channels must be a multiple of 8 and of 4*dconv_comp (LocalState heads)")
        if self.dconv_mode != 1:
            raise UnsupportedConfig("only dconv_mode=1 (DConv in the encoders) is built for HDemucs")
        if self.dconv_attn != self.dconv_lstm or self.dconv_attn != self.norm_starts or self.norm_starts != 4:
            raise UnsupportedConfig("LSTM / attention / GroupNorm must start together at layer 4")
        if self.norm_groups != 4 or self.audio_channels != 2:
            raise UnsupportedConfig("norm_groups=4, audio_channels=2 only")

    @classmethod
    def from_reference_kwargs(cls, *args, **kwargs) -> "HDemucsConfig":
        if args:
            kwargs = dict(kwargs, sources=args[0])
        mine = {}
        for key, value in kwargs.items():
            if key in _PINNED:
                if value != _PINNED[key] and not (key == "multi_freqs" and not value):
                    raise UnsupportedConfig(f"{key}={value!r} is outside the accelerated path (only {_PINNED[key]!r})")
            elif key in _IGNORED:
                continue
            elif key in cls.__dataclass_fields__:
                mine[key] = value
            else:
                raise UnsupportedConfig(f"unknown HDemucs option {key!r}")
        if "sources" in mine:
            mine["sources"] = list(mine["sources"])
        cfg = cls(**mine)
        cfg.validate()
        return cfg

    def reference_kwargs(self) -> dict:
        return dict(sources=list(self.sources), audio_channels=self.audio_channels, channels=self.channels,
                    freq_emb=self.freq_emb, emb_scale=self.emb_scale, norm_starts=self.norm_starts,
                    norm_groups=self.norm_groups, dconv_mode=self.dconv_mode, dconv_depth=self.dconv_depth,
                    dconv_comp=self.dconv_comp, dconv_attn=self.dconv_attn, dconv_lstm=self.dconv_lstm,
                    dconv_init=self.dconv_init, samplerate=self.samplerate, segment=self.segment)


def hdemucs_mmi_config(sources=None) -> HDemucsConfig:
    """``hdemucs_mmi`` (conf/config.yaml:126-165, grids/mmi.py:31, remote/hdemucs_mmi.yaml): 83 637 832 parameters."""
    cfg = HDemucsConfig(sources=list(sources or ["drums", "bass", "other", "vocals"]), channels=48, dconv_comp=4,
                        dconv_init=1e-3, segment=44)
    cfg.validate()
    return cfg


Spec = tp.Tuple[tp.Tuple[int, ...], str, float]


def _dconv_specs(p0: str, ch: int, cfg: HDemucsConfig, lstm: bool, attn: bool, out: dict) -> None:
    hid = int(ch / cfg.dconv_comp)
    for d in range(cfg.dconv_depth):
        p = f"{p0}.dconv.layers.{d}"
        out[f"{p}.0.weight"] = ((hid, ch, 3), "conv", ch * 3)
        out[f"{p}.0.bias"] = ((hid,), "bias", ch * 3)
        out[f"{p}.1.weight"] = ((hid,), "norm_w", 0)
        out[f"{p}.1.bias"] = ((hid,), "norm_b", 0)
        k = 3
        if lstm:      # BLSTM(hid, layers=2, max_steps=200, skip=True), demucs.py:20-67
            for layer, cin in ((0, hid), (1, 2 * hid)):
                for sfx in ("", "_reverse"):
                    out[f"{p}.{k}.lstm.weight_ih_l{layer}{sfx}"] = ((4 * hid, cin), "lstm", hid)
                    out[f"{p}.{k}.lstm.weight_hh_l{layer}{sfx}"] = ((4 * hid, hid), "lstm", hid)
                    out[f"{p}.{k}.lstm.bias_ih_l{layer}{sfx}"] = ((4 * hid,), "lstm", hid)
                    out[f"{p}.{k}.lstm.bias_hh_l{layer}{sfx}"] = ((4 * hid,), "lstm", hid)
            out[f"{p}.{k}.linear.weight"] = ((hid, 2 * hid), "linear", 2 * hid)
            out[f"{p}.{k}.linear.bias"] = ((hid,), "bias", 2 * hid)
            k += 1
        if attn:      # LocalState(hid, heads=4, ndecay=4), demucs.py:157-216
            for n, co in (("content", hid), ("query", hid), ("key", hid), ("query_decay", 16), ("proj", hid)):
                out[f"{p}.{k}.{n}.weight"] = ((co, hid, 1), "decay_w" if n == "query_decay" else "linear", hid)
                out[f"{p}.{k}.{n}.bias"] = ((co,), "decay_b" if n == "query_decay" else "bias", hid)
            k += 1
        out[f"{p}.{k}.weight"] = ((2 * ch, hid, 1), "conv", hid)
        out[f"{p}.{k}.bias"] = ((2 * ch,), "bias", hid)
        out[f"{p}.{k + 1}.weight"] = ((2 * ch,), "norm_w", 0)
        out[f"{p}.{k + 1}.bias"] = ((2 * ch,), "norm_b", 0)
        out[f"{p}.{k + 3}.scale"] = ((ch,), "scale", cfg.dconv_init)


def param_specs(cfg: HDemucsConfig) -> "collections.OrderedDict[str, Spec]":
    """name -> (shape, kind, hint) in the reference's ``state_dict()`` order (encoder, decoder, tencoder, tdecoder,
    freq_emb; 395 tensors / 83 637 832 parameters for hdemucs_mmi)."""
    cfg.validate()
    groups: tp.Dict[str, dict] = {k: {} for k in ("encoder", "decoder", "tencoder", "tdecoder")}
    L = cfg.layers()
    n_time = sum(1 for l in L if l["has_time"])
    for l in L:
        i, j = l["index"], cfg.depth - 1 - l["index"]
        kt = (l["ker"], 1) if l["freq"] else (l["ker"],)
        e = groups["encoder"]
        p = f"encoder.{i}"
        e[f"{p}.conv.weight"] = ((l["chout_z"], l["chin_z"]) + kt, "conv", l["chin_z"] * l["ker"])
        e[f"{p}.conv.bias"] = ((l["chout_z"],), "bias", l["chin_z"] * l["ker"])
        if l["norm"]:
            e[f"{p}.norm1.weight"] = ((l["chout_z"],), "norm_w", 0)
            e[f"{p}.norm1.bias"] = ((l["chout_z"],), "norm_b", 0)
        e[f"{p}.rewrite.weight"] = ((2 * l["chout_z"], l["chout_z"]) + (1,) * len(kt), "conv", l["chout_z"])
        e[f"{p}.rewrite.bias"] = ((2 * l["chout_z"],), "bias", l["chout_z"])
        if l["norm"]:
            e[f"{p}.norm2.weight"] = ((2 * l["chout_z"],), "norm_w", 0)
            e[f"{p}.norm2.bias"] = ((2 * l["chout_z"],), "norm_b", 0)
        _dconv_specs(p, l["chout_z"], cfg, l["lstm"], l["attn"], e)
        d = groups["decoder"]
        p = f"decoder.{j}"
        d[f"{p}.conv_tr.weight"] = ((l["chout_z"], l["dec_out_z"]) + kt, "conv", l["chout_z"] * 2)
        d[f"{p}.conv_tr.bias"] = ((l["dec_out_z"],), "bias", l["chout_z"] * 2)
        if l["norm"]:
            d[f"{p}.norm2.weight"] = ((l["dec_out_z"],), "norm_w", 0)
            d[f"{p}.norm2.bias"] = ((l["dec_out_z"],), "norm_b", 0)
        kr = (3, 3) if l["freq"] else (3,)
        d[f"{p}.rewrite.weight"] = ((2 * l["chout_z"], l["chout_z"]) + kr, "conv", l["chout_z"] * int(math.prod(kr)))
        d[f"{p}.rewrite.bias"] = ((2 * l["chout_z"],), "bias", l["chout_z"] * int(math.prod(kr)))
        if l["norm"]:
            d[f"{p}.norm1.weight"] = ((2 * l["chout_z"],), "norm_w", 0)
            d[f"{p}.norm1.bias"] = ((2 * l["chout_z"],), "norm_b", 0)
        if not l["has_time"]:
            continue
        t = groups["tencoder"]
        p = f"tencoder.{i}"
        t[f"{p}.conv.weight"] = ((l["chout"], l["chin"], 8), "conv", l["chin"] * 8)
        t[f"{p}.conv.bias"] = ((l["chout"],), "bias", l["chin"] * 8)
        if not l["time_empty"]:
            t[f"{p}.rewrite.weight"] = ((2 * l["chout"], l["chout"], 1), "conv", l["chout"])
            t[f"{p}.rewrite.bias"] = ((2 * l["chout"],), "bias", l["chout"])
            _dconv_specs(p, l["chout"], cfg, l["lstm"], l["attn"], t)
        td = groups["tdecoder"]
        p = f"tdecoder.{n_time - 1 - i}"
        td[f"{p}.conv_tr.weight"] = ((l["chout"], l["dec_out"], 8), "conv", l["chout"] * 2)
        td[f"{p}.conv_tr.bias"] = ((l["dec_out"],), "bias", l["chout"] * 2)
        if l["norm"]:
            td[f"{p}.norm2.weight"] = ((l["dec_out"],), "norm_w", 0)
            td[f"{p}.norm2.bias"] = ((l["dec_out"],), "norm_b", 0)
        if not l["time_empty"]:
            td[f"{p}.rewrite.weight"] = ((2 * l["chout"], l["chout"], 3), "conv", l["chout"] * 3)
            td[f"{p}.rewrite.bias"] = ((2 * l["chout"],), "bias", l["chout"] * 3)
    out: "collections.OrderedDict[str, Spec]" = collections.OrderedDict()
    for name in ("encoder", "decoder", "tencoder", "tdecoder"):
        for k, v in sorted(groups[name].items(), key=lambda kv: int(kv[0].split(".")[1])):
            out[k] = v
    if cfg.freq_emb:
        out["freq_emb.embedding.weight"] = ((cfg.nfft // 2 // 4, cfg.channels), "emb", 0)
    return out


def count_params(cfg: HDemucsConfig) -> int:
    return int(sum(int(math.prod(s[0])) for s in param_specs(cfg).values()))


def init_weights(cfg: HDemucsConfig, seed: int = 0, layer_scale: tp.Optional[float] = None):
    """Deterministic synthetic weights (same generator as ``weights.init_weights``; LSTM tensors uniform in
    +-1/sqrt(hidden) as nn.LSTM, the LocalState decay projection as demucs.py:181-184: weight * 0.01, bias -2)."""
    from .weights import init_from_specs
    return init_from_specs(param_specs(cfg), seed, layer_scale, cfg.emb_scale)


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def _register(root: nn.Module, dotted: str, value: torch.Tensor) -> None:
    *path, leaf = dotted.split(".")
    node = root
    for part in path:
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    node.register_parameter(leaf, nn.Parameter(value, requires_grad=False))


class HDemucs(nn.Module):
    """Hybrid Demucs v3, inference only, on the B200 kernel library.  Constructor keywords, parameter names and
    ``forward(mix [B, C, T]) -> [B, S, C, T]`` (any T, hdemucs.py:689-794) follow the reference class; extra keyword-only
    arguments: ``mode`` ("strict" default / "fp32" / ...), ``init_seed`` / ``layer_scale`` (synthetic initialisation)."""

    def __init__(self, sources, *, mode: str = "strict", init_seed: int = 0, layer_scale: tp.Optional[float] = None, **kwargs):
        super().__init__()
        self.cfg = HDemucsConfig.from_reference_kwargs(sources=list(sources), **kwargs)
        self._init_args_kwargs = ((), dict(sources=list(sources), **kwargs))
        cfg = self.cfg
        self.sources = list(cfg.sources)
        self.audio_channels = cfg.audio_channels
        self.samplerate = cfg.samplerate
        self.segment = cfg.segment
        self.nfft, self.hop_length, self.cac, self.depth, self.channels = cfg.nfft, cfg.hop, True, cfg.depth, cfg.channels
        self.hybrid = True
        self.mode = mode
        for name, value in init_weights(cfg, init_seed, layer_scale).items():
            _register(self, name, value)
        self._engines: tp.Dict[tp.Tuple, tp.Any] = {}
        self.train(False)

    @classmethod
    def from_config(cls, cfg: HDemucsConfig, state=None, mode: str = "strict", init_seed: int = 0,
                    layer_scale: tp.Optional[float] = None) -> "HDemucs":
        kw = cfg.reference_kwargs()
        model = cls(kw.pop("sources"), mode=mode, init_seed=init_seed, layer_scale=layer_scale, **kw)
        if state is not None:
            model.load_state_dict(dict(state))
        return model

    @classmethod
    def from_reference(cls, module, mode: str = "strict") -> "HDemucs":
        args, kwargs = module._init_args_kwargs
        kwargs = dict(kwargs)
        if args:
            kwargs["sources"] = args[0]
        model = cls(kwargs.pop("sources"), mode=mode, **kwargs)
        model.load_state_dict(module.state_dict())
        return model

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("demucs_b200.HDemucs is an inference engine; training is out of scope")
        return super().train(False)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engines.clear()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def engine(self):
        from .hdemucs_engine import HDemucsEngine
        p = next(self.parameters())
        key = (p.device, self.mode)
        eng = self._engines.get(key)
        if eng is None:
            eng = HDemucsEngine(self.cfg, dict(self.state_dict()), p.device, self.mode)
            self._engines[key] = eng
        return eng

    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        return self.engine().forward(mix.float())


def hdemucs_mmi(sources=None, **kw) -> HDemucs:
    """The ``hdemucs_mmi`` architecture with synthetic weights (BASELINE configs[1])."""
    cfg = hdemucs_mmi_config(sources)
    return HDemucs.from_config(cfg, **kw)
