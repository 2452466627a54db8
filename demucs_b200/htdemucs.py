"""``HTDemucs`` -- the model object of the drop-in boundary.

Mirrors the reference class of the same name (demucs/htdemucs.py:27-660) as far as inference
callers see it: the constructor keywords, the attributes read by ``apply_model`` /
``Separator`` / ``separate.py`` (``sources, samplerate, audio_channels, segment,
use_train_segment, nfft, hop_length, cac, depth``), ``valid_length`` and
``forward(mix[B, C, T]) -> [B, S, C, T]``.  Parameters carry the reference's names, so
``load_state_dict(reference_model.state_dict())`` works.  The arithmetic is done by
``demucs_b200.engine.Engine`` (hand-written sm_100a kernels); there is no PyTorch fallback.
"""
from __future__ import annotations

from fractions import Fraction
import typing as tp

import torch
from torch import nn

from .config import HTDemucsConfig
from .weights import init_weights, param_specs
from .engine import Engine


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


class HTDemucs(nn.Module):
    """Hybrid Transformer Demucs, inference only, running on the B200 kernel library.

    Args mirror reference ``HTDemucs.__init__`` (htdemucs.py:56-135); options outside the
    released Demucs-v4 space raise ``UnsupportedConfig``.  Extra keyword-only arguments:
      mode: "strict" (default: tcgen05 tensor cores with error-compensated bf16 hi/lo operands, per-stem rel-L2 <= 1e-4
            against the fp32 reference), "bf16" (bf16 tensors and MMAs through the transformer, <= 1e-2), "fp32"
            (CUDA-core FFMA contractions), "tf32x3" / "tf32" (the kind::tf32 forms).
      init_seed / layer_scale: synthetic initialisation (``weights.init_weights``).
    """

    def __init__(self, sources, *, mode: str = "strict", init_seed: int = 0,
                 layer_scale: tp.Optional[float] = None, **kwargs):
        super().__init__()
        self.cfg = HTDemucsConfig.from_reference_kwargs(sources=list(sources), **kwargs)
        self._init_args_kwargs = ((), dict(sources=list(sources), **kwargs))  # states.py:157-163
        cfg = self.cfg
        self.sources = list(cfg.sources)
        self.audio_channels = cfg.audio_channels
        self.samplerate = cfg.samplerate
        self.segment = cfg.segment
        self.use_train_segment = True
        self.nfft = cfg.nfft
        self.hop_length = cfg.hop
        self.cac = True
        self.depth = cfg.depth
        self.channels = cfg.channels
        self.bottom_channels = cfg.bottom_channels
        self.mode = mode
        for name, value in init_weights(cfg, init_seed, layer_scale).items():
            _register(self, name, value)
        self._engines: tp.Dict[tp.Tuple, Engine] = {}
        self.train(False)

    # ---- construction helpers ------------------------------------------------------------
    @classmethod
    def from_reference(cls, module, mode: str = "strict") -> "HTDemucs":
        """Build from a live reference ``demucs.htdemucs.HTDemucs`` (SURVEY.md 8b weight hand-off)."""
        args, kwargs = module._init_args_kwargs
        kwargs = dict(kwargs)
        if args:
            kwargs["sources"] = args[0]
        sources = kwargs.pop("sources")
        model = cls(sources, mode=mode, **kwargs)
        model.load_state_dict(module.state_dict())
        return model

    @classmethod
    def from_config(cls, cfg: HTDemucsConfig, state=None, mode: str = "strict", init_seed: int = 0,
                    layer_scale: tp.Optional[float] = None) -> "HTDemucs":
        kw = cfg.reference_kwargs()
        sources = kw.pop("sources")
        model = cls(sources, mode=mode, init_seed=init_seed, layer_scale=layer_scale, **kw)
        if state is not None:
            model.load_state_dict({k: v for k, v in state.items()})
        return model

    # ---- reference-visible behaviour -------------------------------------------------------
    def valid_length(self, length: int) -> int:
        """htdemucs.py:511-525."""
        training_length = int(self.segment * self.samplerate)
        if training_length < length:
            raise ValueError(f"Given length {length} is longer than training length {training_length}")
        return training_length

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("demucs_b200.HTDemucs is an inference engine; training is out of scope")
        return super().train(False)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engines.clear()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def refresh(self) -> None:
        """Drop the packed device copies; call after modifying parameters in place."""
        self._engines.clear()

    def engine(self) -> Engine:
        """The kernel engine for the parameters' current device (packed weights are cached per
        device, so the reference's habit of moving bag members to the GPU and back,
        apply.py:212-217, does not repack anything)."""
        p = next(self.parameters())
        key = (p.device, self.mode, float(self.segment))
        eng = self._engines.get(key)
        if eng is None:
            cfg = self.cfg
            if self.segment != cfg.segment:
                cfg = HTDemucsConfig(**{**cfg.__dict__, "segment": self.segment})
            state = {k: v for k, v in self.state_dict().items()}
            assert list(state) == list(param_specs(cfg))
            eng = Engine(cfg, state, p.device, self.mode)
            self._engines[key] = eng
        return eng

    def forward(self, mix: torch.Tensor) -> torch.Tensor:
        """mix [B, audio_channels, T <= int(segment*samplerate)] -> [B, S, audio_channels, T]
        (htdemucs.py:527-660).  ``mix`` must live on the same CUDA device as the parameters."""
        return self.engine().forward(mix.float())

    def forward_core(self, mag: torch.Tensor, mix: torch.Tensor) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        """htdemucs.py:662-759 (the ONNX-export surface): mag [B, 2C, F, T] = ``_magnitude(_spec(mix))``,
        mix [B, C, L = training length] -> (spec_out [B, S, 2C, F, T], time_out [B, S, C, L])."""
        return self.engine().forward_core(mag.float(), mix.float())

    def extra_repr(self) -> str:
        return f"sources={self.sources}, segment={float(self.segment):.2f}s, mode={self.mode}"


def htdemucs(sources=None, segment=Fraction(39, 5), **kw) -> HTDemucs:
    """The released ``htdemucs`` architecture with synthetic weights."""
    sources = sources or ["drums", "bass", "other", "vocals"]
    return HTDemucs(sources, segment=segment, dconv_mode=3, bottom_channels=512, **kw)
