"""Parameter inventory and deterministic synthetic initialisation.

Names and shapes are exactly those of the reference module's ``state_dict()``
(533 tensors / 41 984 456 parameters for htdemucs; SURVEY.md section 8b "Weight
hand-off"), so a dict produced here loads into the reference with
``load_state_dict`` and a reference ``state_dict()`` drives the engine unchanged.

Pretrained checkpoints are unreachable offline, so the fixtures are synthetic:
values come from numpy's PCG64 keyed by (seed, parameter name) which is
reproducible on any host, unlike module-construction-order dependent torch init.
"""
from __future__ import annotations

import collections
import typing as tp
import zlib

import numpy as np
import torch

from .config import HTDemucsConfig

Spec = tp.Tuple[tp.Tuple[int, ...], str, float]  # shape, kind, scale hint


def _dconv_specs(prefix: str, ch: int, cfg: HTDemucsConfig, out: dict) -> None:
    hidden = int(ch / cfg.dconv_comp)
    for d in range(cfg.dconv_depth):
        p = f"{prefix}.dconv.layers.{d}"
        out[f"{p}.0.weight"] = ((hidden, ch, 3), "conv", ch * 3)
        out[f"{p}.0.bias"] = ((hidden,), "bias", ch * 3)
        out[f"{p}.1.weight"] = ((hidden,), "norm_w", 0)
        out[f"{p}.1.bias"] = ((hidden,), "norm_b", 0)
        out[f"{p}.3.weight"] = ((2 * ch, hidden, 1), "conv", hidden)
        out[f"{p}.3.bias"] = ((2 * ch,), "bias", hidden)
        out[f"{p}.4.weight"] = ((2 * ch,), "norm_w", 0)
        out[f"{p}.4.bias"] = ((2 * ch,), "norm_b", 0)
        out[f"{p}.6.scale"] = ((ch,), "scale", cfg.dconv_init)


def param_specs(cfg: HTDemucsConfig) -> "collections.OrderedDict[str, Spec]":
    """Ordered name -> (shape, kind, hint) in the reference's registration order
    (htdemucs.py:253-418)."""
    cfg.validate()
    S, A = cfg.n_sources, cfg.audio_channels
    out: "collections.OrderedDict[str, Spec]" = collections.OrderedDict()
    groups: tp.Dict[str, dict] = {k: {} for k in ("encoder", "decoder", "tencoder", "tdecoder")}
    chin_t, chin_z = A, 2 * A
    for i, ch in enumerate(cfg.enc_channels):
        j = cfg.depth - 1 - i  # decoder lists are built with insert(0, ...)
        e, t = groups["encoder"], groups["tencoder"]
        for grp, name, cin, shape_tail in ((e, "encoder", chin_z, (8, 1)), (t, "tencoder", chin_t, (8,))):
            p = f"{name}.{i}"
            grp[f"{p}.conv.weight"] = ((ch, cin) + shape_tail, "conv", cin * 8)
            grp[f"{p}.conv.bias"] = ((ch,), "bias", cin * 8)
            grp[f"{p}.rewrite.weight"] = ((2 * ch, ch) + (1,) * len(shape_tail), "conv", ch)
            grp[f"{p}.rewrite.bias"] = ((2 * ch,), "bias", ch)
            if cfg.dconv_mode & 1:
                _dconv_specs(p, ch, cfg, grp)
        if i == 0:
            chin_t, chin_z = A * S, 2 * A * S
        d, td = groups["decoder"], groups["tdecoder"]
        for grp, name, cout, kt, kr in ((d, "decoder", chin_z, (8, 1), (3, 3)),
                                        (td, "tdecoder", chin_t, (8,), (3,))):
            p = f"{name}.{j}"
            grp[f"{p}.conv_tr.weight"] = ((ch, cout) + kt, "conv", ch * 2)
            grp[f"{p}.conv_tr.bias"] = ((cout,), "bias", ch * 2)
            grp[f"{p}.rewrite.weight"] = ((2 * ch, ch) + kr, "conv", ch * int(np.prod(kr)))
            grp[f"{p}.rewrite.bias"] = ((2 * ch,), "bias", ch * int(np.prod(kr)))
            if cfg.dconv_mode & 2:
                _dconv_specs(p, ch, cfg, grp)
        chin_t = chin_z = ch

    def _sorted(grp):  # state_dict order follows module index
        return sorted(grp.items(), key=lambda kv: (int(kv[0].split(".")[1]),))

    for name in ("encoder", "decoder", "tencoder", "tdecoder"):
        for k, v in _sorted(groups[name]):
            out[k] = v
    if cfg.freq_emb:
        out["freq_emb.embedding.weight"] = ((cfg.nfft // 2 // 4, cfg.channels), "emb", 0)
    cb = cfg.enc_channels[-1]
    if cfg.bottom_channels:
        bc = cfg.bottom_channels
        for name, co, ci in (("channel_upsampler", bc, cb), ("channel_downsampler", cb, bc),
                             ("channel_upsampler_t", bc, cb), ("channel_downsampler_t", cb, bc)):
            out[f"{name}.weight"] = ((co, ci, 1), "conv", ci)
            out[f"{name}.bias"] = ((co,), "bias", ci)
    if cfg.t_layers > 0:
        D, H = cfg.transformer_dim, cfg.ffn_dim
        ct = "crosstransformer"
        for n in ("norm_in", "norm_in_t"):
            out[f"{ct}.{n}.weight"] = ((D,), "norm_w", 0)
            out[f"{ct}.{n}.bias"] = ((D,), "norm_b", 0)
        for branch in ("layers", "layers_t"):
            for i in range(cfg.t_layers):
                p = f"{ct}.{branch}.{i}"
                cross = i % 2 == 1
                attn = "cross_attn" if cross else "self_attn"
                lin = [("linear1", H, D), ("linear2", D, H)]

                def _attn():
                    out[f"{p}.{attn}.in_proj_weight"] = ((3 * D, D), "linear", D)
                    out[f"{p}.{attn}.in_proj_bias"] = ((3 * D,), "bias", D)
                    out[f"{p}.{attn}.out_proj.weight"] = ((D, D), "linear", D)
                    out[f"{p}.{attn}.out_proj.bias"] = ((D,), "bias", D)

                def _lin():
                    for n, co, ci in lin:
                        out[f"{p}.{n}.weight"] = ((co, ci), "linear", ci)
                        out[f"{p}.{n}.bias"] = ((co,), "bias", ci)

                _attn()
                _lin()
                norms = ["norm1", "norm2", "norm3"] if cross else ["norm1", "norm2"]
                for n in norms + ["norm_out"]:
                    out[f"{p}.{n}.weight"] = ((D,), "norm_w", 0)
                    out[f"{p}.{n}.bias"] = ((D,), "norm_b", 0)
                for n in ("gamma_1", "gamma_2"):
                    out[f"{p}.{n}.scale"] = ((D,), "scale", 1e-4)
    return out


def count_params(cfg: HTDemucsConfig) -> int:
    return int(sum(int(np.prod(s[0])) for s in param_specs(cfg).values()))


def init_weights(cfg: HTDemucsConfig, seed: int = 0, layer_scale: tp.Optional[float] = None,
                 dtype=torch.float32) -> "collections.OrderedDict[str, torch.Tensor]":
    """Synthetic weights. ``layer_scale`` overrides every LayerScale value (DConv and
    transformer) -- the reference initialises them at 1e-3 / 1e-4 which hides residual
    branch errors end to end (SURVEY.md section 8c parity caveat); fixtures use 0.5 too.

    Scales follow the reference's effective initial statistics (uniform fan-in init
    followed by ``rescale_module`` for the convolutions, htdemucs.py:365-366), and the
    affine norm parameters are perturbed away from (1, 0) so that they are exercised.
    """
    return init_from_specs(param_specs(cfg), seed, layer_scale, cfg.emb_scale, dtype)


def init_from_specs(specs, seed: int = 0, layer_scale: tp.Optional[float] = None, emb_scale: float = 10.0,
                    dtype=torch.float32) -> "collections.OrderedDict[str, torch.Tensor]":
    """Synthetic values for a name -> (shape, kind, hint) inventory (HTDemucs and HDemucs share the generator)."""
    out = collections.OrderedDict()
    for name, (shape, kind, hint) in specs.items():
        key = zlib.crc32(name.encode()) & 0xFFFFFFFF
        rng = np.random.Generator(np.random.PCG64([seed, key]))
        if kind in ("conv", "linear", "bias"):
            bound = 1.0 / np.sqrt(float(hint))
            if kind != "linear" and not name.startswith("channel_"):
                std = bound / np.sqrt(3.0)
                bound = bound / np.sqrt(std / 0.1)  # rescale_module, demucs.py:69-77
            w = rng.uniform(-bound, bound, size=shape)
        elif kind == "lstm":          # nn.LSTM.reset_parameters: U(-1/sqrt(hidden), 1/sqrt(hidden))
            bound = 1.0 / np.sqrt(float(hint))
            w = rng.uniform(-bound, bound, size=shape)
        elif kind == "decay_w":       # LocalState.query_decay (demucs.py:181-184): weight * 0.01 ...
            w = 0.01 * rng.uniform(-1, 1, size=shape) / np.sqrt(float(hint))
        elif kind == "decay_b":       # ... bias = -2; perturbed so that the four decay rates differ
            w = -2.0 + 0.5 * rng.standard_normal(shape)
        elif kind == "norm_w":
            w = 1.0 + 0.2 * rng.standard_normal(shape)
        elif kind == "norm_b":
            w = 0.1 * rng.standard_normal(shape)
        elif kind == "scale":
            base = hint if layer_scale is None else layer_scale
            w = base * (1.0 + 0.25 * rng.uniform(-1, 1, size=shape))
        elif kind == "emb":
            # smooth cumulative-sum embedding divided by emb_scale (hdemucs.py:52-58)
            w = np.cumsum(rng.standard_normal(shape), axis=0)
            w = w / np.sqrt(np.arange(1, shape[0] + 1))[:, None] / emb_scale
        else:  # pragma: no cover
            raise AssertionError(kind)
        out[name] = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)).to(dtype)
    return out


def check_state_dict(cfg: HTDemucsConfig, state: tp.Mapping[str, torch.Tensor]) -> None:
    """Raise if ``state`` is not a complete, correctly shaped parameter set for ``cfg``."""
    specs = param_specs(cfg)
    missing = [k for k in specs if k not in state]
    extra = [k for k in state if k not in specs]
    if missing or extra:
        raise KeyError(f"state dict mismatch: missing={missing[:4]}... extra={extra[:4]}...")
    for k, (shape, _, _) in specs.items():
        if tuple(state[k].shape) != tuple(shape):
            raise ValueError(f"{k}: expected shape {shape}, got {tuple(state[k].shape)}")
