"""Multi-GPU sharding of ``apply_model`` (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no multi-GPU inference path (its only parallelism is a CPU thread pool,
apply.py:178-182); SURVEY.md section 8e defines the B200-native replacement.  The unit of work
is a segment forward; all units of a (bag member, shift) pass are independent
(apply.py:278-284), so every pass is split into contiguous blocks of segments, one per rank,
with the model weights replicated.  The only data-path exchange is what the overlap-add needs:

* a rank overlap-adds every sample its segments touch except the stretch its FIRST segment shares with the left
  neighbour's last one; that stretch belongs to the left neighbour, who receives just the HEAD of the segment --
  ``seg_len - stride`` samples per row (2.75 MB for htdemucs), not the segment.  Heads are the first thing a rank
  computes and the last thing its neighbour needs, so the transfer hides under the forward passes;
* with several passes (shift trick, bag members) the sample ranges a rank writes move from pass to pass by at
  most the shift; ownership is fixed to the ranges of the first pass and the slivers outside it are sent to
  their owners and added (``combine``) -- no all-reduce of zero-padded full-length tensors;
* what happens to the finished, disjoint ranges is the caller's choice (``gather``): left in place ("none": each
  rank hands its own range to the host, nothing crosses NVLink), collected on rank 0 ("root"), or replicated
  on every rank ("all").

The overlap-add kernel is given the block's global position, so the values are those of a single-GPU run.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


class Shard:
    def __init__(self, group=None, gather: str = "all"):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if gather not in ("all", "root", "none"):
            raise ValueError("gather must be 'all', 'root' or 'none'")
        self.group = group
        self.gather = gather
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.owned: tp.Tuple[int, int] = (0, 0)      # output samples this rank holds after the last apply_model
        self.produced: tp.Tuple[int, int] = (0, 0)   # ... of which it overlap-added these itself (before the gather)

    # ---- partitioning -----------------------------------------------------------------------
    def block_of(self, rank: int, nseg: int) -> tp.Tuple[int, int]:
        """Contiguous block [lo, hi) of segments for ``rank``: sizes differ by at most one,
        larger blocks first (10 min = 103 segments on 8 ranks -> 13,13,13,13,13,13,13,12)."""
        base, extra = divmod(nseg, self.world)
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)

    def block(self, nseg: int) -> tp.Tuple[int, int]:
        return self.block_of(self.rank, nseg)

    @staticmethod
    def halo(seg_len: int, stride: int) -> int:
        """Segments next to a block whose windows reach into it: ceil(seg_len/stride) - 1."""
        return -(-seg_len // stride) - 1

    def _peer(self, rank_in_group: int) -> int:
        return rank_in_group if self.group is None else dist.get_global_rank(self.group, rank_in_group)

    def _comm_device(self, like: torch.Tensor) -> torch.device:
        return like.device if dist.get_backend(self.group) == "nccl" else torch.device("cpu")

    def agree(self, values: tp.List[int]) -> tp.List[int]:
        """Rank 0's integers on every rank (the random shift offsets of apply.py:245: ranks that drew their own
        would cut different windows)."""
        if self.world == 1:
            return list(values)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" \
            else torch.device("cpu")
        t = torch.tensor(list(values), dtype=torch.int64, device=dev)
        dist.broadcast(t, src=self._peer(0), group=self.group)
        return [int(v) for v in t.tolist()]

    # ---- data-path exchange -------------------------------------------------------------------
    def exchange_heads(self, segs: torch.Tensor, lo: int, hi: int, nseg: int, seg_len: int, stride: int, valid: int,
                       length: int) -> tp.Optional[tp.Callable[[], None]]:
        """segs [slots, rows, valid] = [own block [lo, hi) | halo slots for segments hi, hi+1, ...].

        Sends the heads of this rank's first segments to the ranks on the left whose samples they reach, and posts
        the receives for the heads of the segments right of ``hi``.  Every rank derives the same schedule.  Returns a
        function that waits for the transfers and drops the received heads into the halo slots (None: nothing to do).
        """
        if self.world == 1:
            return None
        from .apply import owned_window
        blocks = [self.block_of(r, nseg) for r in range(self.world)]
        halo = self.halo(seg_len, stride)

        def owner(g):
            return next(q for q, (ql, qh) in enumerate(blocks) if ql <= g < qh)

        ops, keep, unpack = [], [], []
        for r, (l, h) in enumerate(blocks):
            if h == l:
                continue                                # empty block: owns no samples, needs no heads
            w1 = owned_window(l, h, nseg, seg_len, stride, length)[1]
            for g in range(h, min(h + halo, nseg)):     # the head of segment g reaches rank r's samples
                n_g = min(length - g * stride, seg_len)
                lead = (valid - n_g) // 2               # centre trim (utils.py:52-53)
                width = min(n_g, w1 - g * stride)
                if width <= 0:
                    continue
                o = owner(g)
                if o == self.rank:
                    buf = segs[g - lo][:, lead:lead + width].contiguous()
                    keep.append(buf)
                    ops.append(dist.P2POp(dist.isend, buf, self._peer(r), self.group))
                elif r == self.rank:
                    buf = torch.empty(segs.shape[1], width, dtype=segs.dtype, device=segs.device)
                    unpack.append(((h - l) + (g - h), lead, width, buf))
                    ops.append(dist.P2POp(dist.irecv, buf, self._peer(o), self.group))
        if not ops:
            return None
        reqs = dist.batch_isend_irecv(ops)

        def finish():
            for req in reqs:
                req.wait()
            for slot, lead, width, buf in unpack:
                segs[slot][:, lead:lead + width].copy_(buf)
            keep.clear()
        return finish

    def pass_range(self, q: int, p: tp.Tuple[int, int, int, int, int], L: int) -> tp.Tuple[int, int]:
        """Output samples [lo, hi) that rank ``q`` overlap-adds in a pass of geometry ``p``."""
        from .apply import owned_window
        nseg, seg_len, stride, length, out_shift = p
        lo, hi = self.block_of(q, nseg)
        w0, w1 = owned_window(lo, hi, nseg, seg_len, stride, length)
        return min(max(w0 - out_shift, 0), L), min(max(w1 - out_shift, 0), L)

    def sliver_schedule(self, passes, L: int):
        """(own, sched): ownership = the ranges of the first pass (they tile [0, L)); sched = the pieces
        (sender q, owner r, lo, hi) of what later passes wrote outside the sender's own range.  Pure arithmetic:
        every rank derives the same schedule."""
        own = [self.pass_range(q, passes[0], L) for q in range(self.world)]
        sched = []
        for q in range(self.world):
            spans = [self.pass_range(q, p, L) for p in passes]
            spans = [(a, b) for a, b in spans if b > a]
            if not spans:
                continue
            lo_q, hi_q = min(a for a, _ in spans), max(b for _, b in spans)
            if own[q][1] <= own[q][0]:          # owns nothing: everything it wrote belongs to somebody else
                outside = [(lo_q, hi_q)]
            else:
                outside = [(lo_q, min(own[q][0], hi_q)), (max(own[q][1], lo_q), hi_q)]
            for a, b in outside:
                for r in range(self.world):
                    x, y = max(a, own[r][0]), min(b, own[r][1])
                    if r != q and y > x:
                        sched.append((q, r, x, y))
        return own, sched

    def combine(self, out: torch.Tensor, passes: tp.List[tp.Tuple[int, int, int, int, int]]) -> tp.Tuple[int, int]:
        """``out`` [rows, L] holds what this rank overlap-added in every pass; ``passes`` lists each pass's
        (nseg, seg_len, stride, window length, out_shift), from which every rank derives every rank's sample ranges.
        Ownership = the ranges of the first pass (they tile [0, L)); contributions a later pass wrote outside them are
        sent to their owners and added.  Then the ``gather`` policy.  Returns the range of ``out`` that is valid here.
        """
        L = out.shape[-1]
        if self.world == 1:
            self.owned = self.produced = (0, L)
            return self.owned

        own, sched = self.sliver_schedule(passes, L)
        ops, keep, adds = [], [], []
        for q, r, x, y in sched:
            if q == self.rank:
                buf = out[:, x:y].contiguous()
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, self._peer(r), self.group))
            elif r == self.rank:
                buf = torch.empty(out.shape[0], y - x, dtype=out.dtype, device=out.device)
                adds.append((x, y, buf))
                ops.append(dist.P2POp(dist.irecv, buf, self._peer(q), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            for x, y, buf in adds:
                out[:, x:y] += buf
        a, b = own[self.rank]
        self.owned = self.produced = (a, b)
        if self.gather == "none":
            return self.owned
        rows = out.shape[0]
        if self.gather == "all":
            width = max(y - x for x, y in own)
            local = torch.empty(rows, width, dtype=out.dtype, device=out.device)
            local[:, :b - a].copy_(out[:, a:b])
            gathered = torch.empty(self.world * rows, width, dtype=out.dtype, device=out.device)   # rank-major
            dist.all_gather_into_tensor(gathered, local, group=self.group)
            for q, (x, y) in enumerate(own):
                if q != self.rank and y > x:
                    out[:, x:y].copy_(gathered[q * rows:(q + 1) * rows, :y - x])
            self.owned = (0, L)
            return self.owned
        # "root": the pieces travel to rank 0 only
        ops, recvs, keep = [], [], []
        for q, (x, y) in enumerate(own):
            if q == 0 or y <= x:
                continue
            if self.rank == q:
                buf = out[:, x:y].contiguous()
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, self._peer(0), self.group))
            elif self.rank == 0:
                buf = torch.empty(rows, y - x, dtype=out.dtype, device=out.device)
                recvs.append((x, y, buf))
                ops.append(dist.P2POp(dist.irecv, buf, self._peer(q), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            for x, y, buf in recvs:
                out[:, x:y].copy_(buf)
        if self.rank == 0:
            self.owned = (0, L)
        return self.owned
