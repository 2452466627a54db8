"""``apply_model`` -- device-resident segment batcher (drop-in for reference demucs/apply.py).

Same signature, result and side effects as the reference's ``apply_model``
(apply.py:145-322) for HTDemucs models and ``BagOfModels`` of them, but organised for a GPU:

* the whole track moves to the compute device once (the reference copies every padded chunk H2D and every
  weighted chunk D2H, apply.py:295,312); a host result streams back range by range on a copy stream, under the
  separation of the later segments;
* the segments of a pass are cut out of the track ``batch_size`` at a time by one gather kernel
  (``bd_gather_segments``) and the engine writes its result straight into the pass's segment store;
* overlap-add, centre trim, un-shift, shift averaging and bag weighting happen in one gather kernel per batch
  (K8, csrc/ola.cu) over the samples that batch completed, instead of one read-modify-write of the track per
  segment;
* with ``torch.distributed`` initialised and ``group`` given, the segments of every pass are
  sharded across ranks (demucs_b200/distributed.py).

Python's global ``random`` stream is consumed exactly as the reference does -- one
``random.randint`` per shift (apply.py:245) and one ``random.randrange(1)`` per segment forward
(transformer.py:680 via htdemucs.py:593) -- so a seeded run draws the same shift offsets.
"""
from __future__ import annotations

import contextlib
import copy
import random
from threading import Lock
import typing as tp

import torch
from torch import nn

from . import _lib
from ._lib import ptr
from .hdemucs import HDemucs
from .htdemucs import HTDemucs

Model = tp.Union[HTDemucs, HDemucs]


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
            if segment is not None:       # apply.py:53-55: overrides the segment of non-HT members, in place
                if not isinstance(other, HTDemucs) and segment > other.segment:
                    other.segment = segment
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
    if not isinstance(model, HTDemucs):
        # no ``valid_length``: every chunk is evaluated at its own length (apply.py:302-309), so the store is as wide as
        # the widest chunk and a shorter last chunk is run on its own
        seg_len = min(seg_len, max(length, 1)) if not split else seg_len
        return min(seg_len, length), seg_len, stride, offsets
    # leaf: HTDemucs with an explicit segment pads to it, otherwise to the training length
    valid = int(segment * model.samplerate) if segment is not None else train_len
    if min(seg_len, length) > valid or valid > train_len:
        raise ValueError(f"Given length {max(min(seg_len, length), valid)} is longer than "
                         f"training length {train_len}")
    return valid, min(seg_len, max(length, 1)) if not split else seg_len, stride, offsets


def owned_window(lo_seg: int, hi_seg: int, nseg: int, seg_len: int, stride: int, length: int) -> tp.Tuple[int, int]:
    """Window samples [w0, w1) that the holder of segments [lo_seg, hi_seg) overlap-adds: everything its segments
    touch EXCEPT the part of its first segments that the left neighbour's last segment also reaches -- that part
    belongs to the left neighbour, who receives the heads of these segments (distributed.Shard.exchange_heads).
    Boundaries are block edges shifted right by the overlap seg_len - stride, so the ranges of consecutive blocks tile
    the window."""
    if hi_seg <= lo_seg:
        return 0, 0
    reach = max(seg_len - stride, 0)
    w0 = 0 if lo_seg == 0 else min(length, lo_seg * stride + reach)
    w1 = length if hi_seg >= nseg else min(length, hi_seg * stride + reach)
    return w0, max(w0, w1)


def run_pass(model: HTDemucs, track: torch.Tensor, ps: _Pass, out: torch.Tensor, row_alpha, accumulate: bool,
             split: bool, overlap: float, transition_power: float, segment, batch_size: int,
             notify=None, progress_bar=None, shard=None, sink=None) -> tp.Tuple[int, int, int, int, int]:
    """Separate one window of ``track`` [B, C, Ltrack] and overlap-add it into ``out`` [B*S*C, L].

    Segments go through the engine ``batch_size`` at a time: one gather kernel cuts the batch out of the track, the
    forward writes straight into the pass's segment store, and one overlap-add launch finishes the window samples
    that no later segment can touch; ``sink(lo, hi)`` is told which output samples just became final.
    ``shard`` (``distributed.Shard``) restricts this rank to a contiguous block of the segments; the heads of the
    right neighbours' first segments arrive through ``shard.exchange_heads``.  Returns the pass geometry
    (nseg, seg_len, stride, window length, out_shift) from which ``Shard.combine`` derives who wrote what.
    """
    eng = model.engine()
    B, Cc, Ltrack = track.shape
    S = len(model.sources)
    rows = B * S * Cc
    valid, seg_len, stride, offsets = _segment_plan(model, ps.length, split, overlap, segment)
    nseg = len(offsets)
    weight = transition_weight(seg_len, transition_power, track.device) if split else \
        torch.ones(seg_len, device=track.device)
    halo = -(-seg_len // stride) - 1            # later segments whose heads reach back into a block's samples
    if shard is None:
        lo_seg, hi_seg = 0, nseg
    else:
        lo_seg, hi_seg = shard.block(nseg)
    n_halo = min(halo, nseg - hi_seg) if hi_seg > lo_seg else 0
    n_slots = max(hi_seg - lo_seg + n_halo, 1)
    key = ("apply", rows, valid)
    segs = eng._buf(key, "segs", n_slots * rows * valid).view(n_slots, rows, valid)
    fixed = isinstance(model, HTDemucs)          # HTDemucs pads every chunk to one length; v3 runs chunks as they are
    if shard is not None and not fixed:
        raise NotImplementedError("sharding across ranks is built for HTDemucs models (fixed-length segments); run v3 "
                                  "models unsharded, or shard the tracks of a batch across ranks yourself")
    if fixed and model.cfg.t_layers > 0:
        # the reference draws random.randrange(1) inside every segment forward (transformer.py:680); every
        # rank draws for ALL segments so that sharded ranks keep identical RNG streams
        for _ in range(nseg):
            random.randrange(1)
    w0, w1 = owned_window(lo_seg, hi_seg, nseg, seg_len, stride, ps.length)
    out_len = out.shape[-1]

    def to_out(n):      # window sample -> output sample, clipped
        return min(max(n - ps.out_shift, 0), out_len)

    pending = None
    done = w0           # window samples [w0, done) have been overlap-added
    sent = shard is None
    for s0 in range(lo_seg, hi_seg, batch_size):
        s1 = min(s0 + batch_size, hi_seg)
        if notify:
            for i in range(s0, s1):
                notify(offsets[i], "start")
        # v3 (no valid_length): the trailing chunks that the window cuts short run one by one at their own length
        n = s1 - s0 if fixed else sum(1 for i in range(s0, s1) if ps.length - i * stride >= valid)
        if n > 0:
            batch = eng._buf(key, "batch", n * B * Cc * valid).view(n * B, Cc, valid)
            eng._k("bd_gather_segments", ptr(track), ptr(batch), B, Cc, Ltrack, ps.offset0, ps.length, s0, n, seg_len, stride,
                   valid, eng._stream(), nbytes=8.0 * n * B * Cc * valid)
            eng.forward(batch, out=segs[s0 - lo_seg: s0 - lo_seg + n])       # [n*B, S, C, valid] = [n, rows, valid]
        for i in range(s0 + n, s1):
            # evaluated at its own length and centred in its slot, where the overlap-add's centre trim looks for it
            n_i = ps.length - i * stride
            lead = (valid - n_i) // 2
            one = eng._buf(key, "batch_short", B * Cc * n_i).view(B, Cc, n_i)
            eng._k("bd_gather_segments", ptr(track), ptr(one), B, Cc, Ltrack, ps.offset0 + i * stride, n_i, 0, 1,
                   n_i, n_i, n_i, eng._stream(), nbytes=8.0 * B * Cc * n_i)
            segs[i - lo_seg][:, lead:lead + n_i].copy_(eng.forward(one).view(rows, n_i))
        if notify:
            for i in range(s0, s1):
                notify(offsets[i], "end")
        if progress_bar is not None:
            progress_bar.update(s1 - s0)
        if not sent and s1 >= min(hi_seg, lo_seg + halo):
            # the heads of this block's first segments are ready: ship them to the left neighbours, and post the
            # receives for the heads this block needs from the right
            pending = shard.exchange_heads(segs, lo_seg, hi_seg, nseg, seg_len, stride, valid, ps.length)
            sent = True
        # samples below s1*stride are final unless this was the block's last batch (then the heads decide)
        upto = min(w1, s1 * stride) if s1 < hi_seg else None
        if upto is not None and upto > done:
            _overlap_add(eng, segs, weight, out, lo_seg, s1 - lo_seg, nseg, rows, valid, seg_len, stride, ps, done, upto,
                         row_alpha, accumulate)
            if sink is not None:
                sink(to_out(done), to_out(upto))
            done = upto
    if shard is not None and not sent:
        pending = shard.exchange_heads(segs, lo_seg, hi_seg, nseg, seg_len, stride, valid, ps.length)
    if pending is not None:
        pending()                                                      # heads have landed in the halo slots
    if w1 > done:
        _overlap_add(eng, segs, weight, out, lo_seg, hi_seg - lo_seg + n_halo, nseg, rows, valid, seg_len, stride, ps, done,
                     w1, row_alpha, accumulate)
        if sink is not None:
            sink(to_out(done), to_out(w1))
    return nseg, seg_len, stride, ps.length, ps.out_shift


def _overlap_add(eng, segs, weight, out, seg_first, nseg_local, nseg, rows, valid, seg_len, stride, ps, n_begin, n_end,
                 row_alpha, accumulate) -> None:
    eng._k("bd_overlap_add", ptr(segs), ptr(weight), ptr(out), seg_first, nseg_local, nseg, rows, valid, seg_len, stride,
           ps.length, out.shape[-1], ps.out_shift, n_begin, n_end, ptr(row_alpha), ps.alpha, int(accumulate),
           eng._stream(), nbytes=4.0 * rows * (n_end - n_begin) * (2 + (1 if accumulate else 0)))


LAST_IO = {"h2d_bytes": 0, "d2h_bytes": 0}     # host <-> device bytes of the last apply_model call on host buffers


class _HostSink:
    """Streams finished output ranges to a pinned host tensor on a side stream while later segments compute."""

    def __init__(self, dev_out: torch.Tensor, device: torch.device):
        self.dev_out = dev_out
        self.host = torch.empty(dev_out.shape, dtype=dev_out.dtype, pin_memory=True)
        self.stream = torch.cuda.Stream(device)
        self.device = device
        self.bytes = 0

    def __call__(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        ev = torch.cuda.current_stream(self.device).record_event()
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            for r in range(self.dev_out.shape[0]):   # contiguous device run -> contiguous pinned run: plain async copies
                self.host[r, lo:hi].copy_(self.dev_out[r, lo:hi], non_blocking=True)
        self.bytes += self.dev_out.shape[0] * (hi - lo) * 4

    def finish(self) -> torch.Tensor:
        self.stream.synchronize()
        return self.host


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

    A host ``mix`` is uploaded once (asynchronously when it is pinned) and the stems come back in a pinned host
    tensor, streamed range by range on a copy stream while later segments are still being separated.
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
    on_cuda = device.type == "cuda"
    guard = torch.cuda.device(device) if on_cuda else contextlib.nullcontext()
    with guard:
        to_host = on_cuda and mix.device.type == "cpu"
        if to_host and shard is not None and not shifts and mix.dtype == torch.float32:
            # a rank uploads only the stretch of the track its own segments read (centre padding included)
            valid, seg_len, stride, offsets = _segment_plan(models[0], length, split, overlap, segment)
            lo_seg, hi_seg = shard.block(len(offsets))
            starts = [i * stride - (valid - min(length - i * stride, seg_len)) // 2 for i in range(lo_seg, hi_seg)]
            lo = max(0, min(starts, default=0))
            hi = min(length, max(starts, default=0) + valid) if starts else 0
            track = torch.empty(batch, channels, length, dtype=torch.float32, device=device)
            for b_ in range(batch if hi > lo else 0):
                for c_ in range(channels):     # contiguous pinned run -> contiguous device run
                    track[b_, c_, lo:hi].copy_(mix[b_, c_, lo:hi], non_blocking=mix.is_pinned())
            h2d_bytes = batch * channels * max(hi - lo, 0) * 4
        else:
            track = mix.to(device=device, dtype=torch.float32, non_blocking=to_host and mix.is_pinned())   # one H2D
            h2d_bytes = mix.numel() * 4 if to_host else 0
        n_passes = len(models) * max(shifts, 1)
        # every pass but the first accumulates; a sharded multi-pass run also accumulates into ranges its first pass
        # did not write (the owned ranges move with the shift), so only that case needs a zero fill
        alloc = torch.zeros if (shard is not None and n_passes > 1) else torch.empty
        out = alloc(batch * S * channels, length, device=device)
        totals = [0.] * S
        if bag_weights is not None:
            for w_m in bag_weights:
                for k, w in enumerate(w_m):
                    totals[k] += w
        bar = None
        sink = _HostSink(out, device) if (to_host and n_passes == 1) else None
        written: tp.List[tp.Tuple[int, int]] = []
        n_done = 0
        for mi, sub in enumerate(models):
            original_device = next(iter(sub.parameters())).device
            sub.to(device)
            sub.eval()
            if bag_weights is not None:
                # estimates += w[m][k] * out_m[:, k]; afterwards /= totals[k]  (apply.py:219-228), folded into the
                # per-row factor of the overlap-add
                ra = torch.tensor([w / t for w, t in zip(bag_weights[mi], totals)], dtype=torch.float32)
                row_alpha = ra.view(1, S, 1).expand(batch, S, channels).reshape(-1).contiguous().to(device)
            else:
                row_alpha = None
            src = track
            if shifts:
                max_shift = int(0.5 * sub.samplerate)
                src = tensor_chunk(track).padded(length + 2 * max_shift)
            for si in range(max(shifts, 1)):
                if shifts:
                    # drawn where the reference draws it (apply.py:245), between the segment forwards of consecutive
                    # shifts (which consume the stream too); sharded ranks must cut the same windows: rank 0's draw wins
                    offset = random.randint(0, max_shift)
                    if shard is not None:
                        offset = shard.agree([offset])[0]
                    ps = _Pass(mi, si, offset, length + max_shift - offset, max_shift - offset, 1.0 / shifts)
                else:
                    ps = _Pass(mi, 0, 0, length, 0, 1.0)

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
                written.append(run_pass(sub, src, ps, out, row_alpha, accumulate=n_done > 0, split=split, overlap=overlap,
                                        transition_power=transition_power, segment=segment, batch_size=batch_size,
                                        notify=notify if callback is not None else None, progress_bar=bar, shard=shard,
                                        sink=sink))
                n_done += 1
            sub.to(original_device)
        if bar is not None:
            bar.close()
        own = (0, length)
        if shard is not None:
            own = shard.combine(out, written)          # slivers to their owners, then gather as the shard is set up
        if not to_host:
            return out.view(batch, S, channels, length).to(mix.device)
        if sink is None:
            sink = _HostSink(out, device)
            sink(*own)
        elif shard is not None:
            # the range this rank produced went to the host while it was being made; what a gather policy added
            # around it ("all", or "root" on rank 0) follows now
            sink(own[0], shard.produced[0])
            sink(shard.produced[1], own[1])
        LAST_IO["h2d_bytes"], LAST_IO["d2h_bytes"] = h2d_bytes, sink.bytes
        return sink.finish().view(batch, S, channels, length)
