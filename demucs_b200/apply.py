"""``apply_model`` -- device-resident segment batcher (drop-in for reference demucs/apply.py).

Same signature, result and side effects as the reference's ``apply_model``
(apply.py:145-322) for HTDemucs models and ``BagOfModels`` of them, but organised for a GPU:

* the whole track moves to the compute device once and the separated stems move back once
  (the reference copies every padded chunk H2D and every weighted chunk D2H, apply.py:295,312);
* the work units (bag member, shift, segment) are enumerated up front, the segments of a pass
  are cut into one batch tensor and run through the kernel engine ``batch_size`` at a time;
* overlap-add, centre trim, un-shift, shift averaging and bag weighting happen in ONE gather
  kernel per pass (K8, csrc/ola.cu) instead of one read-modify-write of the track per segment;
* with ``torch.distributed`` initialised and ``group`` given, the segments of every pass are
  sharded across ranks (demucs_b200/distributed.py).

Python's global ``random`` stream is consumed exactly as the reference does -- one
``random.randint`` per shift (apply.py:245) and one ``random.randrange(1)`` per segment forward
(transformer.py:680 via htdemucs.py:593) -- so a seeded run draws the same shift offsets.
"""
from __future__ import annotations

import copy
import random
from threading import Lock
import typing as tp

import torch
from torch import nn

from . import _lib
from ._lib import ptr
from .htdemucs import HTDemucs

Model = HTDemucs


class BagOfModels(nn.Module):
    """Reference apply.py:29-79: models sharing sources/samplerate/channels with per-model,
    per-source weights.  ``forward`` is not callable; use ``apply_model``."""

    def __init__(self, models: tp.List[Model], weights: tp.Optional[tp.List[tp.List[float]]] = None,
                 segment: tp.Optional[float] = None):
        super().__init__()
        assert len(models) > 0
        first = models[0]
        for other in models:
            assert other.sources == first.sources
            assert other.samplerate == first.samplerate
            assert other.audio_channels == first.audio_channels
            # `segment` only overrides non-HT models in the reference (apply.py:54-56); every model
            # here is an HTDemucs, whose segment is bound to its training length.
        self.audio_channels = first.audio_channels
        self.samplerate = first.samplerate
        self.sources = first.sources
        self.models = nn.ModuleList(models)
        if weights is None:
            weights = [[1. for _ in first.sources] for _ in models]
        else:
            assert len(weights) == len(models)
            for weight in weights:
                assert len(weight) == len(first.sources)
        self.weights = weights

    @property
    def max_allowed_segment(self) -> float:
        out = float('inf')
        for model in self.models:
            if isinstance(model, HTDemucs):
                out = min(out, float(model.segment))
        return out

    def forward(self, x):
        raise NotImplementedError("Call `apply_model` on this.")


class TensorChunk:
    """Lazy window on a tensor with centred zero padding (reference apply.py:82-124)."""

    def __init__(self, tensor, offset=0, length=None):
        total_length = tensor.shape[-1]
        assert offset >= 0
        assert offset < total_length
        length = total_length - offset if length is None else min(total_length - offset, length)
        if isinstance(tensor, TensorChunk):
            self.tensor = tensor.tensor
            self.offset = offset + tensor.offset
        else:
            self.tensor = tensor
            self.offset = offset
        self.length = length
        self.device = tensor.device

    @property
    def shape(self):
        shape = list(self.tensor.shape)
        shape[-1] = self.length
        return shape

    def window(self, target_length: int) -> tp.Tuple[int, int, int, int]:
        """(lo, hi, pad_left, pad_right): the slice of the parent and the zeros around it that
        make up ``padded(target_length)``."""
        delta = target_length - self.length
        total_length = self.tensor.shape[-1]
        assert delta >= 0
        start = self.offset - delta // 2
        end = start + target_length
        lo, hi = max(0, start), min(total_length, end)
        return lo, hi, lo - start, end - hi

    def padded(self, target_length: int) -> torch.Tensor:
        lo, hi, left, right = self.window(target_length)
        out = torch.nn.functional.pad(self.tensor[..., lo:hi], (left, right))
        assert out.shape[-1] == target_length
        return out


def tensor_chunk(tensor_or_chunk):
    if isinstance(tensor_or_chunk, TensorChunk):
        return tensor_or_chunk
    assert isinstance(tensor_or_chunk, torch.Tensor)
    return TensorChunk(tensor_or_chunk)


def center_trim(tensor: torch.Tensor, reference: tp.Union[torch.Tensor, int]):
    """Reference utils.py:38-54."""
    ref_size = reference.size(-1) if isinstance(reference, torch.Tensor) else reference
    delta = tensor.size(-1) - ref_size
    if delta < 0:
        raise ValueError("tensor must be larger than reference. " f"Delta is {delta}.")
    if delta:
        tensor = tensor[..., delta // 2:-(delta - delta // 2)]
    return tensor


def _replace_dict(_dict: tp.Optional[dict], *subs: tp.Tuple[tp.Hashable, tp.Any]) -> dict:
    _dict = {} if _dict is None else copy.copy(_dict)
    for key, value in subs:
        _dict[key] = value
    return _dict


def transition_weight(segment_length: int, transition_power: float, device) -> torch.Tensor:
    """Triangular weight of apply.py:271-276, computed with the same torch ops."""
    weight = torch.cat([torch.arange(1, segment_length // 2 + 1, device=device),
                        torch.arange(segment_length - segment_length // 2, 0, -1, device=device)])
    assert len(weight) == segment_length
    return ((weight / weight.max()) ** transition_power).float().contiguous()


class _Pass(tp.NamedTuple):
    """One (bag member, shift) sweep over a window of the (padded) track."""
    model_idx: int
    shift_idx: int
    offset0: int        # start of the window in `track`
    length: int         # window length
    out_shift: int      # first window sample that lands in the output (max_shift - offset)
    alpha: float        # 1 / shifts


def _segment_plan(model: HTDemucs, length: int, split: bool, overlap: float,
                  segment: tp.Optional[float]) -> tp.Tuple[int, int, int, tp.List[int]]:
    """(valid_length, seg_len, stride, offsets) of the split / leaf branches (apply.py:257-284,302-312)."""
    train_len = int(model.segment * model.samplerate)
    if split:
        seg = model.segment if segment is None else segment
        assert seg is not None and seg > 0.
        seg_len = int(model.samplerate * seg)
        stride = int((1 - overlap) * seg_len)
        offsets = list(range(0, length, stride))
    else:
        seg_len, stride, offsets = length, length, [0]
    # leaf: HTDemucs with an explicit segment pads to it, otherwise to the training length
    valid = int(segment * model.samplerate) if segment is not None else train_len
    if min(seg_len, length) > valid or valid > train_len:
        raise ValueError(f"Given length {max(min(seg_len, length), valid)} is longer than "
                         f"training length {train_len}")
    return valid, min(seg_len, max(length, 1)) if not split else seg_len, stride, offsets


def run_pass(model: HTDemucs, track: torch.Tensor, ps: _Pass, out: torch.Tensor, row_alpha, accumulate: bool,
             split: bool, overlap: float, transition_power: float, segment, batch_size: int,
             notify=None, progress_bar=None, shard=None) -> None:
    """Separate one window of ``track`` [B, C, Ltrack] and overlap-add it into ``out`` [B*S*C, L].

    ``shard`` (``distributed.Shard``) restricts this rank to a contiguous block of the segments;
    the left neighbour's trailing segments arrive through ``shard.exchange_halo``.
    """
    eng = model.engine()
    B, Cc, _ = track.shape
    S = len(model.sources)
    rows = B * S * Cc
    valid, seg_len, stride, offsets = _segment_plan(model, ps.length, split, overlap, segment)
    nseg = len(offsets)
    weight = transition_weight(seg_len, transition_power, track.device) if split else \
        torch.ones(seg_len, device=track.device)
    if shard is None:
        lo_seg, hi_seg, first = 0, nseg, 0
    else:
        lo_seg, hi_seg = shard.block(nseg)
        first = max(0, lo_seg - shard.halo(seg_len, stride))   # segments reaching into this rank's samples
    n_local = hi_seg - first
    key = ("apply", rows, valid)
    segs = eng._buf(key, "segs", max(n_local, 1) * rows * valid).view(max(n_local, 1), rows, valid)
    parent = TensorChunk(track, ps.offset0, ps.length)
    if model.cfg.t_layers > 0:
        # the reference draws random.randrange(1) inside every segment forward (transformer.py:680); every
        # rank draws for ALL segments so that sharded ranks keep identical RNG streams
        for _ in range(nseg):
            random.randrange(1)
    for s0 in range(lo_seg, hi_seg, batch_size):
        idx = list(range(s0, min(s0 + batch_size, hi_seg)))
        batch = eng._buf(key, "batch", len(idx) * B * Cc * valid).view(len(idx) * B, Cc, valid)
        batch.zero_()
        for j, i in enumerate(idx):
            lo, hi, left, _ = TensorChunk(parent, offsets[i], seg_len).window(valid)
            batch[j * B:(j + 1) * B, :, left:left + hi - lo].copy_(track[..., lo:hi])
        for i in idx:
            if notify:
                notify(offsets[i], "start")
        res = eng.forward(batch)                                   # [n*B, S, C, valid]
        # [n, B, S*C, valid] -> rows ordered (b, s, c) per segment
        segs[s0 - first: s0 - first + len(idx)].copy_(res.view(len(idx), rows, valid))
        if notify:
            for i in idx:
                notify(offsets[i], "end")
        if progress_bar is not None:
            progress_bar.update(len(idx))
    n_begin, n_end = 0, ps.length
    if shard is not None:
        shard.exchange_halo(segs, lo_seg - first, lo_seg, hi_seg, nseg, shard.halo(seg_len, stride))
        n_begin = lo_seg * stride
        n_end = ps.length if hi_seg >= nseg else hi_seg * stride
        if n_local <= 0:
            return nseg, stride, ps.length
    eng._k("bd_overlap_add", ptr(segs), ptr(weight), ptr(out), first, n_local, nseg, rows, valid, seg_len, stride,
           ps.length, out.shape[-1], ps.out_shift, n_begin, n_end, ptr(row_alpha), ps.alpha, int(accumulate),
           eng._stream(), nbytes=4.0 * rows * (n_local * min(seg_len, ps.length) + (n_end - n_begin) * (2 if accumulate else 1)))
    return nseg, stride, ps.length


def apply_model(model: tp.Union[BagOfModels, Model],
                mix: tp.Union[torch.Tensor, TensorChunk],
                shifts: int = 1, split: bool = True,
                overlap: float = 0.25, transition_power: float = 1.,
                progress: bool = False, device=None,
                num_workers: int = 0, segment: tp.Optional[float] = None,
                pool=None, lock=None,
                callback: tp.Optional[tp.Callable[[dict], None]] = None,
                callback_arg: tp.Optional[dict] = None,
                batch_size: int = 16, shard=None) -> torch.Tensor:
    """Apply model to a given mixture -- same contract as reference apply.py:145-322.

    mix [B, C, L] (any device) -> [B, S, C, L] on ``mix.device``; computation on ``device``
    (default ``mix.device``, as in the reference; it must be a CUDA device -- there is no CPU path).
    ``num_workers`` / ``pool`` are accepted for signature compatibility; on a GPU the reference
    ignores them too (apply.py:178-182).  Extra: ``batch_size`` segments per forward, ``shard``
    (``distributed.Shard``) to split every pass across ranks.
    """
    if isinstance(mix, TensorChunk):
        mix = mix.padded(mix.length)
    device = mix.device if device is None else torch.device(device)
    if lock is None:
        lock = Lock()
    callback_arg = _replace_dict(callback_arg, *{"model_idx_in_bag": 0, "shift_idx": 0, "segment_offset": 0}.items())
    assert transition_power >= 1, "transition_power < 1 leads to weird behavior."
    if isinstance(model, BagOfModels):
        models, bag_weights = list(model.models), model.weights
    else:
        models, bag_weights = [model], None
    callback_arg["models"] = len(models)
    batch, channels, length = mix.shape
    S = len(models[0].sources)
    track = mix.to(device=device, dtype=torch.float32)             # one H2D for the whole track
    out = torch.zeros(batch * S * channels, length, device=device)
    totals = [0.] * S
    bar = None
    plans: tp.Set[tp.Optional[tp.Tuple[int, int, int]]] = set()     # segment plans of the passes (None: shifted)
    for mi, sub in enumerate(models):
        original_device = next(iter(sub.parameters())).device
        sub.to(device)
        sub.eval()
        if bag_weights is not None:
            # estimates += w[m][k] * out_m[:, k]; afterwards /= totals[k]  (apply.py:219-228)
            for k, w in enumerate(bag_weights[mi]):
                totals[k] += w
            ra = torch.tensor(bag_weights[mi], dtype=torch.float32).view(1, S, 1).expand(batch, S, channels)
            row_alpha = ra.reshape(-1).contiguous().to(device)
        else:
            row_alpha = None
        passes: tp.List[_Pass] = []
        src = track
        if shifts:
            max_shift = int(0.5 * sub.samplerate)
            src = tensor_chunk(track).padded(length + 2 * max_shift)
        for si in range(max(shifts, 1)):
            if shifts:
                offset = random.randint(0, max_shift)
                ps = _Pass(mi, si, offset, length + max_shift - offset, max_shift - offset, 1.0 / shifts)
            else:
                ps = _Pass(mi, 0, 0, length, 0, 1.0)
            passes.append(ps)

            def notify(seg_offset, state, ps=ps):
                if callback is not None:
                    with lock:
                        callback(_replace_dict(callback_arg, ("model_idx_in_bag", ps.model_idx),
                                               ("shift_idx", ps.shift_idx), ("segment_offset", seg_offset),
                                               ("state", state)))

            if progress and split and bar is None:
                import tqdm
                seg_s = float(sub.segment if segment is None else segment)
                scale = float(format((1 - overlap) * seg_s, ".2f"))
                bar = tqdm.tqdm(unit_scale=scale, ncols=120, unit='seconds')
            plan = run_pass(sub, src, ps, out, row_alpha, accumulate=(mi > 0 or si > 0 or shard is not None),
                     split=split, overlap=overlap,
                     transition_power=transition_power, segment=segment, batch_size=batch_size,
                     notify=notify if callback is not None else None, progress_bar=bar, shard=shard)
            plans.add(plan if not shifts else None)
        sub.to(original_device)
    if bar is not None:
        bar.close()
    if shard is not None:
        # ranks hold disjoint sample ranges of every pass: one all-gather when all passes share one unshifted
        # segment plan, else an all-reduce of the zero-padded pieces
        only = next(iter(plans)) if len(plans) == 1 else None
        shard.combine(out, only)
    out = out.view(batch, S, channels, length)
    if bag_weights is not None:
        out /= torch.tensor(totals, dtype=torch.float32, device=device).view(1, S, 1, 1)
    return out.to(mix.device)                                      # one D2H for all stems
