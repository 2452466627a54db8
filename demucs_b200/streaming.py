"""Streaming separation: segment by segment with carried overlap-add state.

The reference processes whole tracks (``apply_model``, apply.py:257-301); its web front end (``web/``) is the use case
for something incremental.  ``StreamSeparator`` accepts audio as it arrives and returns every stretch of the stems as
soon as no later segment can change it -- after segment k has been separated, the samples below ``(k+1)*stride`` are
final -- so the latency is one segment (7.8 s of audio for htdemucs) plus one forward, and the memory is bounded: the
input kept is one segment plus the left context the last, centre-padded chunk may need (apply.py:108-124), the state
carried between calls is the last ``ceil(seg/stride) - 1`` separated segments whose tails still overlap what comes next.

The result equals ``apply_model(model, whole_track, shifts=0, split=True, overlap=overlap)`` on the concatenated input:
the same gather (``bd_gather_segments``), forward and overlap-add (``bd_overlap_add``) kernels run on the same windows;
only the bookkeeping differs.  HTDemucs models only (fixed segment length).
"""
from __future__ import annotations

import random
import typing as tp

import torch

from ._lib import ptr
from .apply import transition_weight
from .htdemucs import HTDemucs


class StreamSeparator:
    def __init__(self, model: HTDemucs, overlap: float = 0.25, transition_power: float = 1.0, device=None):
        if not isinstance(model, HTDemucs):
            raise TypeError("StreamSeparator needs an HTDemucs model (a fixed segment length)")
        assert transition_power >= 1, "transition_power < 1 leads to weird behavior."
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        model.to(self.device)
        model.eval()
        self.eng = model.engine()
        self.seg_len = int(model.samplerate * model.segment)           # apply.py:263-264
        self.stride = int((1 - overlap) * self.seg_len)
        self.halo = -(-self.seg_len // self.stride) - 1
        self.channels = model.audio_channels
        self.rows = len(model.sources) * self.channels
        self.weight = transition_weight(self.seg_len, transition_power, self.device)
        self.reset()

    def reset(self) -> None:
        self.pending = torch.empty(1, self.channels, 0, device=self.device)    # input from global sample `base` on
        self.base = 0
        self.received = 0
        self.next_seg = 0            # first segment not separated yet
        self.emitted = 0             # output samples handed out so far
        self.kept = 0                # separated segments held in `store` (the last ones, in order)
        self.store = None

    # ------------------------------------------------------------------------------------------------------------
    def push(self, chunk: torch.Tensor) -> torch.Tensor:
        """chunk [C, n] (or [1, C, n]) of new audio -> stems [S, C, m] that just became final (m may be 0)."""
        if chunk.dim() == 2:
            chunk = chunk[None]
        if chunk.dim() != 3 or chunk.shape[0] != 1 or chunk.shape[1] != self.channels:
            raise ValueError(f"expected a chunk of shape [{self.channels}, n]")
        self.pending = torch.cat([self.pending, chunk.to(device=self.device, dtype=torch.float32)], dim=-1)
        self.received += chunk.shape[-1]
        out = []
        while self.next_seg * self.stride + self.seg_len <= self.received:      # a whole segment has arrived
            out.append(self._separate(self.next_seg, total=None))
        return self._cat(out)

    def flush(self) -> torch.Tensor:
        """End of the stream: the remaining segments (the last ones cut short and centre-padded as apply.py:108-124)."""
        out = []
        total = self.received
        while self.next_seg * self.stride < total:
            out.append(self._separate(self.next_seg, total=total))
        res = self._cat(out)
        self.reset()
        return res

    # ------------------------------------------------------------------------------------------------------------
    def _cat(self, parts: tp.List[torch.Tensor]) -> torch.Tensor:
        S = len(self.model.sources)
        if not parts:
            return torch.empty(S, self.channels, 0, device=self.device)
        return torch.cat(parts, dim=-1)

    def _separate(self, k: int, total: tp.Optional[int]) -> torch.Tensor:
        eng, seg_len, stride, rows = self.eng, self.seg_len, self.stride, self.rows
        valid = seg_len
        # while streaming the track has no end yet: any length that leaves segment k whole tiles the same way
        length = total if total is not None else k * stride + seg_len
        nseg = -(-length // stride)
        if self.model.cfg.t_layers > 0:
            random.randrange(1)                                   # transformer.py:680, as apply_model consumes it
        slots = self.halo + 1
        if self.store is None:
            self.store = torch.empty(slots, rows, valid, device=self.device)
        if self.kept == slots:                                    # the oldest segment no longer reaches new samples
            self.store[:-1].copy_(self.store[1:].clone())
            self.kept -= 1
        with torch.cuda.device(self.device) if self.device.type == "cuda" else _Null():
            batch = eng._buf(("stream", rows, valid), "batch", self.channels * valid).view(1, self.channels, valid)
            eng._k("bd_gather_segments", ptr(self.pending), ptr(batch), 1, self.channels, self.pending.shape[-1], -self.base,
                   length, k, 1, seg_len, stride, valid, eng._stream(), nbytes=8.0 * self.channels * valid)
            eng.forward(batch, out=self.store[self.kept: self.kept + 1])
            self.kept += 1
            n_begin = k * stride
            n_end = min((k + 1) * stride, length) if (total is None or (k + 1) * stride < total) else length
            out = torch.empty(rows, n_end - n_begin, device=self.device)
            eng._k("bd_overlap_add", ptr(self.store), ptr(self.weight), ptr(out), k - self.kept + 1, self.kept, nseg, rows, valid,
                   seg_len, stride, length, n_end - n_begin, n_begin, n_begin, n_end, None, 1.0, 0, eng._stream(),
                   nbytes=4.0 * rows * (n_end - n_begin) * 3)
        self.next_seg = k + 1
        self.emitted = n_end
        # input older than what the next segment (or its centre padding, at most valid/2 of left context) can read goes
        keep_from = max(0, (k + 1) * stride - valid // 2 - 1)
        if keep_from > self.base:
            self.pending = self.pending[..., keep_from - self.base:].contiguous()
            self.base = keep_from
        return out.view(len(self.model.sources), self.channels, n_end - n_begin)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
