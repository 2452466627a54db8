"""Architecture description of the HTDemucs separation path.

The reference carries its hyper-parameters as ``HTDemucs.__init__`` keyword
arguments (reference ``demucs/htdemucs.py:56-135``) recorded by ``capture_init``
(``demucs/states.py:157-163``).  The engine only implements the slice of that
space the released Demucs-v4 models live in (complex-as-channels, no GroupNorm in
the outer layers, sinusoidal embeddings, dense attention); anything else is
rejected loudly here instead of being silently mis-computed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction
import math
import typing as tp


class UnsupportedConfig(ValueError):
    """Raised when reference kwargs fall outside of the accelerated path."""


# reference kwargs that must keep their released value (htdemucs.py:56-135)
_PINNED = {
    "wiener_iters": 0, "end_iters": 0, "wiener_residual": False, "cac": True,
    "rewrite": True, "multi_freqs": None, "kernel_size": 8, "stride": 4,
    "context": 1, "context_enc": 0, "t_emb": "sin",
    "t_norm_in": True, "t_norm_in_group": False, "t_group_norm": False,
    "t_norm_first": True, "t_norm_out": True, "t_layer_scale": True,
    "t_gelu": True, "t_sin_random_shift": 0, "t_sparse_self_attn": False,
    "t_sparse_cross_attn": False, "t_cross_first": False, "channels_time": None,
    "growth": 2, "nfft": 4096, "depth": 4, "use_train_segment": True,
}
# kwargs that only matter for training / initialisation and are ignored here
_IGNORED = {
    "multi_freqs_depth", "emb_smooth", "time_stride", "norm_groups", "t_max_positions",
    "t_weight_decay", "t_lr", "t_cape_mean_normalize", "t_cape_augment",
    "t_cape_glob_loc_scale", "t_mask_type", "t_mask_random_seed", "t_sparse_attn_window",
    "t_global_window", "t_sparsity", "t_auto_sparsity", "rescale", "dconv_attn", "dconv_lstm",
    "t_dropout",          # inert in eval mode; the released htdemucs models record 0.02 (grids/mmi.py:22)
}


@dataclass
class HTDemucsConfig:
    sources: tp.List[str] = field(default_factory=lambda: ["drums", "bass", "other", "vocals"])
    audio_channels: int = 2
    channels: int = 48
    nfft: int = 4096
    depth: int = 4
    freq_emb: float = 0.2
    emb_scale: float = 10.0
    norm_starts: int = 4
    dconv_mode: int = 1
    dconv_depth: int = 2
    dconv_comp: int = 8
    dconv_init: float = 1e-3
    bottom_channels: int = 0
    t_layers: int = 5
    t_hidden_scale: float = 4.0
    t_heads: int = 8
    t_max_period: float = 10000.0
    t_weight_pos_embed: float = 1.0
    samplerate: int = 44100
    segment: tp.Union[Fraction, float, int] = 10

    # ---- derived quantities -------------------------------------------------
    @property
    def hop(self) -> int:
        return self.nfft // 4

    @property
    def n_sources(self) -> int:
        return len(self.sources)

    @property
    def segment_length(self) -> int:
        """``int(segment * samplerate)`` as in ``htdemucs.py:519``."""
        return int(self.segment * self.samplerate)

    @property
    def enc_channels(self) -> tp.List[int]:
        return [self.channels * 2 ** i for i in range(self.depth)]

    @property
    def transformer_dim(self) -> int:
        return self.bottom_channels or self.enc_channels[-1]

    @property
    def ffn_dim(self) -> int:
        return int(self.transformer_dim * self.t_hidden_scale)

    def frames(self, length: int) -> int:
        return int(math.ceil(length / self.hop))

    def time_lengths(self, length: int) -> tp.List[int]:
        """Lengths of the time branch after each encoder layer (hdemucs.py:131-135)."""
        out = [length]
        for _ in range(self.depth):
            out.append(int(math.ceil(out[-1] / 4)))
        return out

    def validate(self) -> None:
        if self.nfft != 4096 or self.depth != 4:
            raise UnsupportedConfig("only nfft=4096, depth=4 is built (the Demucs-v4 geometry)")
        if self.norm_starts < self.depth:
            raise UnsupportedConfig("GroupNorm inside HEnc/HDecLayer (norm_starts < depth) is not built")
        if self.channels % self.dconv_comp or self.channels % 8:
            raise UnsupportedConfig("channels must be a multiple of 8 and of dconv_comp")
        dim = self.transformer_dim
        if self.t_layers > 0 and (dim % self.t_heads or dim // self.t_heads != 64):
            raise UnsupportedConfig("attention kernels are built for head_dim == 64")
        if self.t_layers > 0 and dim % 4:
            raise UnsupportedConfig("transformer dim must be a multiple of 4")
        if self.dconv_mode not in (0, 1, 2, 3):
            raise UnsupportedConfig("dconv_mode must be in 0..3")

    # ---- construction from reference kwargs ------------------------------------
    @classmethod
    def from_reference_kwargs(cls, *args, **kwargs) -> "HTDemucsConfig":
        """Build from the (args, kwargs) recorded by the reference's ``capture_init``."""
        if args:
            kwargs = dict(kwargs)
            kwargs["sources"] = args[0]
            if len(args) > 1:
                raise UnsupportedConfig("pass HTDemucs options by keyword")
        mine = {}
        for key, value in kwargs.items():
            if key in _PINNED:
                if value != _PINNED[key]:
                    raise UnsupportedConfig(f"{key}={value!r} is outside the accelerated path "
                                            f"(only {_PINNED[key]!r})")
            elif key in _IGNORED:
                continue
            elif key in cls.__dataclass_fields__:
                mine[key] = value
            else:
                raise UnsupportedConfig(f"unknown HTDemucs option {key!r}")
        if "sources" in mine:
            mine["sources"] = list(mine["sources"])
        cfg = cls(**mine)
        cfg.validate()
        return cfg

    def reference_kwargs(self) -> dict:
        """kwargs that build the same architecture with the reference's HTDemucs."""
        return dict(sources=list(self.sources), audio_channels=self.audio_channels,
                    channels=self.channels, freq_emb=self.freq_emb, emb_scale=self.emb_scale,
                    norm_starts=self.norm_starts, dconv_mode=self.dconv_mode,
                    dconv_depth=self.dconv_depth, dconv_comp=self.dconv_comp,
                    dconv_init=self.dconv_init, bottom_channels=self.bottom_channels,
                    t_layers=self.t_layers, t_hidden_scale=self.t_hidden_scale,
                    t_heads=self.t_heads, t_max_period=self.t_max_period,
                    t_weight_pos_embed=self.t_weight_pos_embed, samplerate=self.samplerate,
                    segment=self.segment)


def htdemucs_config(sources=None, segment=Fraction(39, 5)) -> HTDemucsConfig:
    """The released ``htdemucs`` architecture (conf/config.yaml:195-271 +
    grids/mmi.py:15-30): dconv_mode=3, bottom_channels=512, 5 transformer layers."""
    cfg = HTDemucsConfig(sources=list(sources or ["drums", "bass", "other", "vocals"]),
                         dconv_mode=3, bottom_channels=512, segment=segment)
    cfg.validate()
    return cfg


def htdemucs_6s_config(segment=Fraction(39, 5)) -> HTDemucsConfig:
    return htdemucs_config(["drums", "bass", "other", "vocals", "guitar", "piano"], segment)
