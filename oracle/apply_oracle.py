"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by ``demucs_b200``.

Flat (non-recursive) CPU restatement of the reference's ``apply_model``
(demucs/apply.py:145-322) for HTDemucs models and bags of them, written as an
explicit enumeration of (bag member, shift, segment) work units.  Pinned against the
reference by tests/test_oracle_vs_reference.py and the golden fixtures.
"""
from __future__ import annotations

import random
import typing as tp

import torch

from .htdemucs_oracle import htdemucs_forward


def padded_chunk(track: torch.Tensor, offset: int, length: int, target: int) -> torch.Tensor:
    """``TensorChunk(track, offset, length).padded(target)`` (apply.py:82-124): the window is
    centred, and whatever falls outside the parent tensor is zero (real neighbouring audio
    is used when it exists)."""
    total = track.shape[-1]
    length = min(total - offset, length)
    delta = target - length
    start = offset - delta // 2
    end = start + target
    lo, hi = max(0, start), min(total, end)
    return torch.nn.functional.pad(track[..., lo:hi], (lo - start, end - hi))


def center_trim(x: torch.Tensor, length: int) -> torch.Tensor:
    """utils.py:38-54 -- odd surplus drops the extra sample on the right."""
    delta = x.shape[-1] - length
    if delta < 0:
        raise ValueError("tensor must be larger than reference")
    return x[..., delta // 2: x.shape[-1] - (delta - delta // 2)] if delta else x


def transition_weight(segment_length: int, power: float, dtype) -> torch.Tensor:
    """Triangular segment weight (apply.py:271-276)."""
    half = segment_length // 2
    w = torch.cat([torch.arange(1, half + 1), torch.arange(segment_length - half, 0, -1)]).to(dtype)
    return (w / w.max()) ** power


def apply_single(W, cfg, mix: torch.Tensor, shifts: int, split: bool, overlap: float,
                 transition_power: float, segment: tp.Optional[float],
                 consume_rng: bool = True, window: tp.Optional[tp.Tuple[int, int]] = None) -> torch.Tensor:
    """shifts / split / leaf branches of apply.py:231-322 for one HTDemucs model.

    ``consume_rng``: the reference draws ``random.randrange(1)`` inside every forward
    (transformer.py:680, sin_random_shift=0) which advances Python's global RNG between the
    per-shift ``random.randint`` draws (apply.py:245); reproduced to keep shift offsets equal.

    ``window`` = (lo, hi): only output samples [lo, hi) are wanted (a spot check on a long track): segments that do
    not touch them are not evaluated (their RNG draws still happen); the result is exact inside the window and
    meaningless outside.
    """
    assert transition_power >= 1
    B, C, L = mix.shape
    S = cfg.n_sources
    is_ht = hasattr(cfg, "t_layers")       # HTDemucs pads to its training length; HDemucs v3 has no valid_length
    if not is_ht:
        from .hdemucs_oracle import hdemucs_forward

    def leaf(track, offset, length):
        """forward on one centred, zero-padded window of ``track`` (apply.py:302-322)."""
        if not is_ht:                     # apply.py:306-309: valid_length = length
            return hdemucs_forward(W, cfg, padded_chunk(track, offset, length, min(track.shape[-1] - offset, length)))
        valid = int(segment * cfg.samplerate) if segment is not None else cfg.segment_length
        if valid < length:
            raise ValueError(f"Given length {length} is longer than training length {valid}")
        if consume_rng and cfg.t_layers > 0:
            random.randrange(1)
        out = htdemucs_forward(W, cfg, padded_chunk(track, offset, length, valid))
        return center_trim(out, length)

    def split_pass(track, offset0, length, out_shift=0):
        """split branch (apply.py:257-301) over the window [offset0, offset0+length) of track."""
        if not split:
            return leaf(track, offset0, length)
        seg = segment if segment is not None else cfg.segment
        seg_len = int(cfg.samplerate * seg)
        stride = int((1 - overlap) * seg_len)
        weight = transition_weight(seg_len, transition_power, mix.dtype)
        out = mix.new_zeros(B, S, C, length)
        sumw = mix.new_zeros(length)
        for off in range(0, length, stride):
            n = min(length - off, seg_len)
            # nested TensorChunk: offsets add and the length clips to the window, but padding
            # is cut from the underlying tensor (apply.py:87-96,108-124)
            sumw[off: off + n] += weight[:n]
            if window is not None and (off + n - out_shift <= window[0] or off - out_shift >= window[1]):
                if consume_rng and is_ht and cfg.t_layers > 0:
                    random.randrange(1)
                continue
            chunk = leaf(track, offset0 + off, n)
            out[..., off: off + n] += weight[:n] * chunk
        assert sumw.min() > 0
        return out / sumw

    if shifts:
        max_shift = int(0.5 * cfg.samplerate)
        padded = padded_chunk(mix, 0, L, L + 2 * max_shift)
        acc = mix.new_zeros(B, S, C, L)
        for _ in range(shifts):
            offset = random.randint(0, max_shift)
            res = split_pass(padded, offset, L + max_shift - offset, max_shift - offset)
            acc += res[..., max_shift - offset:]
        return acc / shifts
    return split_pass(mix, 0, L)


def apply_model_oracle(models, mix: torch.Tensor, shifts: int = 1, split: bool = True,
                       overlap: float = 0.25, transition_power: float = 1.0,
                       segment: tp.Optional[float] = None,
                       bag_weights: tp.Optional[tp.List[tp.List[float]]] = None,
                       window: tp.Optional[tp.Tuple[int, int]] = None) -> torch.Tensor:
    """``models``: a (weights, cfg) pair or a list of them (a bag, apply.py:201-229).  ``window``: see apply_single."""
    if isinstance(models, tuple):
        W, cfg = models
        return apply_single(W, cfg, mix, shifts, split, overlap, transition_power, segment, window=window)
    S = models[0][1].n_sources
    if bag_weights is None:
        bag_weights = [[1.0] * S for _ in models]
    totals = [0.0] * S
    est = 0.0
    for (W, cfg), mw in zip(models, bag_weights):
        out = apply_single(W, cfg, mix, shifts, split, overlap, transition_power, segment, window=window)
        for k, w in enumerate(mw):
            out[:, k] *= w
            totals[k] += w
        est = est + out
    for k in range(S):
        est[:, k] /= totals[k]
    return est
