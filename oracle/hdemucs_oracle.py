"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by ``demucs_b200``.

Functional CPU restatement of ``HDemucs.forward`` (reference demucs/hdemucs.py:689-794) for the hdemucs_mmi family
(cac, hybrid, no MultiWrap): HEncLayer / HDecLayer with GroupNorm, ``inject`` and ``empty`` (hdemucs.py:123-157,
304-335), DConv with BLSTM and LocalState (demucs.py:20-67,86-216), written over a weight dict with explicit loops for
the recurrence and the frame splitting.  Pinned against the unmodified reference by tests/golden/hdemucs_*.npz
(oracle/make_golden.py: hdemucs_fixture) and live by tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import math
import typing as tp

import torch
import torch.nn.functional as F

from .htdemucs_oracle import stft_cac, istft_cac

Weights = tp.Mapping[str, torch.Tensor]


def lstm_direction(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, reverse: bool) -> torch.Tensor:
    """One direction of one nn.LSTM layer: x [T, N, in] -> [T, N, h]; gate order i, f, g, o."""
    T, N, _ = x.shape
    h = x.new_zeros(N, w_hh.shape[1])
    c = x.new_zeros(N, w_hh.shape[1])
    pre = x @ w_ih.t() + (b_ih + b_hh)
    out = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        gates = pre[t] + h @ w_hh.t()
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out)


def blstm(x: torch.Tensor, W: Weights, p: str, max_steps: int = 200) -> torch.Tensor:
    """BLSTM(dim, layers=2, max_steps=200, skip=True).forward (demucs.py:37-67) on [B, C, T]."""
    B, C, T = x.shape
    y = x
    framed = T > max_steps
    if framed:
        width, stride = max_steps, max_steps // 2
        nframes = math.ceil(T / stride)                                  # utils.unfold (utils.py:20-35)
        xp = F.pad(x, (0, (nframes - 1) * stride + width - T))
        frames = torch.stack([xp[..., k * stride: k * stride + width] for k in range(nframes)], dim=2)   # [B,C,F,K]
        x = frames.permute(0, 2, 1, 3).reshape(-1, C, width)
    s = x.permute(2, 0, 1)                                               # [T, N, C]
    for layer in range(2):
        outs = []
        for sfx in ("", "_reverse"):
            outs.append(lstm_direction(s, W[f"{p}.lstm.weight_ih_l{layer}{sfx}"], W[f"{p}.lstm.weight_hh_l{layer}{sfx}"],
                                       W[f"{p}.lstm.bias_ih_l{layer}{sfx}"], W[f"{p}.lstm.bias_hh_l{layer}{sfx}"],
                                       reverse=bool(sfx)))
        s = torch.cat(outs, dim=-1)
    s = s @ W[f"{p}.linear.weight"].t() + W[f"{p}.linear.bias"]
    x = s.permute(1, 2, 0)
    if framed:
        fr = x.reshape(B, -1, C, width)
        limit = stride // 2
        out = []
        for k in range(nframes):
            if k == 0:
                out.append(fr[:, k, :, :-limit])
            elif k == nframes - 1:
                out.append(fr[:, k, :, limit:])
            else:
                out.append(fr[:, k, :, limit:-limit])
        x = torch.cat(out, -1)[..., :T]
    return x + y


def local_state(x: torch.Tensor, W: Weights, p: str, heads: int = 4, ndecay: int = 4) -> torch.Tensor:
    """LocalState(channels, heads=4, nfreqs=0, ndecay=4).forward (demucs.py:186-216) on [B, C, T]."""
    B, C, T = x.shape
    idx = torch.arange(T, dtype=x.dtype)
    delta = idx[:, None] - idx[None, :]                                   # keys t (rows) - queries s (columns)
    q = F.conv1d(x, W[f"{p}.query.weight"], W[f"{p}.query.bias"]).view(B, heads, -1, T)
    k = F.conv1d(x, W[f"{p}.key.weight"], W[f"{p}.key.bias"]).view(B, heads, -1, T)
    dots = torch.einsum("bhct,bhcs->bhts", k, q) / math.sqrt(k.shape[2])
    decays = torch.arange(1, ndecay + 1, dtype=x.dtype)
    dq = torch.sigmoid(F.conv1d(x, W[f"{p}.query_decay.weight"], W[f"{p}.query_decay.bias"]).view(B, heads, -1, T)) / 2
    kernel = -decays.view(-1, 1, 1) * delta.abs() / math.sqrt(ndecay)
    dots = dots + torch.einsum("fts,bhfs->bhts", kernel, dq)
    dots = dots.masked_fill(torch.eye(T, dtype=torch.bool), -100)
    w = torch.softmax(dots, dim=2)
    content = F.conv1d(x, W[f"{p}.content.weight"], W[f"{p}.content.bias"]).view(B, heads, -1, T)
    res = torch.einsum("bhts,bhct->bhcs", w, content).reshape(B, -1, T)
    return x + F.conv1d(res, W[f"{p}.proj.weight"], W[f"{p}.proj.bias"])


def dconv(x: torch.Tensor, W: Weights, prefix: str, depth: int, lstm: bool, attn: bool, taps=None) -> torch.Tensor:
    """DConv (demucs.py:86-154) on [N, C, T]; module indices shift when BLSTM / LocalState are inserted (:146-149)."""
    for d in range(depth):
        p = f"{prefix}.layers.{d}"
        dil = 2 ** d
        h = F.conv1d(x, W[f"{p}.0.weight"], W[f"{p}.0.bias"], dilation=dil, padding=dil)
        h = F.gelu(F.group_norm(h, 1, W[f"{p}.1.weight"], W[f"{p}.1.bias"], 1e-5))
        k = 3
        if lstm:
            h = blstm(h, W, f"{p}.{k}")
            if taps is not None:
                taps[f"{p}.lstm"] = h
            k += 1
        if attn:
            h = local_state(h, W, f"{p}.{k}")
            if taps is not None:
                taps[f"{p}.attn"] = h
            k += 1
        u = F.conv1d(h, W[f"{p}.{k}.weight"], W[f"{p}.{k}.bias"])
        u = F.glu(F.group_norm(u, 1, W[f"{p}.{k + 1}.weight"], W[f"{p}.{k + 1}.bias"], 1e-5), dim=1)
        x = x + W[f"{p}.{k + 3}.scale"][:, None] * u
    return x


def _norm(x, W, name, on: bool, groups: int):
    return F.group_norm(x, groups, W[f"{name}.weight"], W[f"{name}.bias"], 1e-5) if on else x


def enc_layer(x, W, prefix: str, l: dict, cfg, freq: bool, empty: bool = False, inject=None, taps=None):
    """HEncLayer.forward (hdemucs.py:123-157)."""
    if not freq and x.dim() == 4:
        x = x.view(x.shape[0], -1, x.shape[-1])
    ker, stride = (l["ker"], l["stride"]) if (freq or not l["freq"]) else (8, 4)
    pad = ker // 4 if (l["pad"] or not freq) else 0
    if freq:
        y = F.conv2d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=(stride, 1), padding=(pad, 0))
    else:
        le = x.shape[-1]
        if le % stride:
            x = F.pad(x, (0, stride - le % stride))
        y = F.conv1d(x, W[f"{prefix}.conv.weight"], W[f"{prefix}.conv.bias"], stride=stride, padding=pad)
    if empty:
        return y
    if inject is not None:
        y = y + (inject[:, :, None] if (inject.dim() == 3 and y.dim() == 4) else inject)
    y = F.gelu(_norm(y, W, f"{prefix}.norm1", l["norm"], cfg.norm_groups))
    if freq:
        B, C, Fr, T = y.shape
        y = dconv(y.permute(0, 2, 1, 3).reshape(-1, C, T), W, f"{prefix}.dconv", cfg.dconv_depth, l["lstm"], l["attn"], taps)
        y = y.view(B, Fr, C, T).permute(0, 2, 1, 3)
        z = F.conv2d(y, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"])
    else:
        y = dconv(y, W, f"{prefix}.dconv", cfg.dconv_depth, l["lstm"], l["attn"], taps)
        z = F.conv1d(y, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"])
    return F.glu(_norm(z, W, f"{prefix}.norm2", l["norm"], cfg.norm_groups), dim=1)


def dec_layer(x, skip, length: int, W, prefix: str, l: dict, cfg, freq: bool, last: bool, empty: bool = False):
    """HDecLayer.forward (hdemucs.py:304-335) -> (z, pre)."""
    if freq and x.dim() == 3:
        x = x.view(x.shape[0], l["chout_z"], -1, x.shape[-1])
    ker, stride = (l["ker"], l["stride"]) if (freq or not l["freq"]) else (8, 4)
    pad = ker // 4 if (l["pad"] or not freq) else 0
    if not empty:
        x = x + skip
        if freq:
            y = F.conv2d(x, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=1)
        else:
            y = F.conv1d(x, W[f"{prefix}.rewrite.weight"], W[f"{prefix}.rewrite.bias"], padding=1)
        y = F.glu(_norm(y, W, f"{prefix}.norm1", l["norm"], cfg.norm_groups), dim=1)
    else:
        y = x
    if freq:
        z = F.conv_transpose2d(y, W[f"{prefix}.conv_tr.weight"], W[f"{prefix}.conv_tr.bias"], stride=(stride, 1))
        z = _norm(z, W, f"{prefix}.norm2", l["norm"], cfg.norm_groups)
        if pad:
            z = z[..., pad:-pad, :]
    else:
        z = F.conv_transpose1d(y, W[f"{prefix}.conv_tr.weight"], W[f"{prefix}.conv_tr.bias"], stride=stride)
        z = _norm(z, W, f"{prefix}.norm2", l["norm"], cfg.norm_groups)
        z = z[..., pad: pad + length]
    return (z if last else F.gelu(z)), y


def hdemucs_forward(W: Weights, cfg, mix: torch.Tensor, taps: tp.Optional[dict] = None) -> torch.Tensor:
    """HDemucs.forward in eval mode (hdemucs.py:689-794): mix [B, C, L] -> [B, S, C, L] (any L)."""
    dt = mix.dtype
    W = {k: v.to(dt) for k, v in W.items()}
    B, C, L = mix.shape
    mag = stft_cac(mix, cfg.nfft)
    if taps is not None:
        taps["stft"] = mag
    mean, std = mag.mean(dim=(1, 2, 3), keepdim=True), mag.std(dim=(1, 2, 3), keepdim=True)
    x = (mag - mean) / (1e-5 + std)
    meant, stdt = mix.mean(dim=(1, 2), keepdim=True), mix.std(dim=(1, 2), keepdim=True)
    xt = (mix - meant) / (1e-5 + stdt)
    layers = cfg.layers()
    n_time = sum(1 for l in layers if l["has_time"])
    saved, saved_t, lengths, lengths_t = [], [], [], []
    for l in layers:
        i = l["index"]
        lengths.append(x.shape[-1])
        inject = None
        if l["has_time"]:
            lengths_t.append(xt.shape[-1])
            xt = enc_layer(xt, W, f"tencoder.{i}", l, cfg, freq=False, empty=l["time_empty"])
            if l["time_empty"]:
                inject = xt
            else:
                saved_t.append(xt)
            if taps is not None:
                taps[f"tenc{i}"] = xt
        x = enc_layer(x, W, f"encoder.{i}", l, cfg, freq=l["freq"], inject=inject, taps=taps)
        if i == 0 and cfg.freq_emb:
            x = x + cfg.freq_emb * (cfg.emb_scale * W["freq_emb.embedding.weight"]).t()[None, :, :, None]
        saved.append(x)
        if taps is not None:
            taps[f"enc{i}"] = x
    x = torch.zeros_like(x)
    xt = torch.zeros_like(x)
    offset = cfg.depth - n_time
    for j in range(cfg.depth):
        l = layers[cfg.depth - 1 - j]
        x, pre = dec_layer(x, saved.pop(), lengths.pop(), W, f"decoder.{j}", l, cfg, freq=l["freq"], last=j == cfg.depth - 1)
        if taps is not None:
            taps[f"dec{j}"] = x
        if j >= offset:
            length_t = lengths_t.pop()
            if l["time_empty"]:
                assert pre.shape[2] == 1
                xt, _ = dec_layer(pre[:, :, 0], None, length_t, W, f"tdecoder.{j - offset}", l, cfg, freq=False, last=False,
                                  empty=True)
            else:
                xt, _ = dec_layer(xt, saved_t.pop(), length_t, W, f"tdecoder.{j - offset}", l, cfg, freq=False,
                                  last=j == cfg.depth - 1)
            if taps is not None:
                taps[f"tdec{j - offset}"] = xt
    S = cfg.n_sources
    Fq, T = x.shape[-2:]
    x = x.view(B, S, -1, Fq, T) * std[:, None] + mean[:, None]
    x = istft_cac(x, L, cfg.nfft)
    xt = xt.view(B, S, -1, L) * stdt[:, None] + meant[:, None]
    if taps is not None:
        taps["istft"], taps["time_out"] = x, xt
    return xt + x
