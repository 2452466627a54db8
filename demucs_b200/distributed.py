"""Multi-GPU sharding of ``apply_model`` (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no multi-GPU inference path (its only parallelism is a CPU thread pool,
apply.py:178-182); SURVEY.md section 8e defines the B200-native replacement.  The unit of work
is a segment forward; all units of a (bag member, shift) pass are independent
(apply.py:278-284), so every pass is split into contiguous blocks of segments, one per rank,
with the model weights replicated.  The only data-path exchange is what the overlap-add
needs: rank r owns the output samples [lo_r*stride, hi_r*stride) and those also receive the
tails of the ``halo`` segments just left of its block, which the left neighbour sends
(point-to-point, ~11 MB per segment).  Afterwards every rank holds a disjoint range of the
result; ``combine`` exchanges the pieces with one all-gather (or, when the shift trick moves the
owned ranges from pass to pass, sums the zero-padded pieces with one all-reduce).  The overlap-add kernel is given the block's global position, so the values
are bit-identical to a single-GPU run.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


class Shard:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

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
        """Segments to the left of a block whose windows reach into it: ceil(seg_len/stride) - 1."""
        return -(-seg_len // stride) - 1

    # ---- data-path exchange -------------------------------------------------------------------
    def exchange_halo(self, segs: torch.Tensor, n_halo: int, lo: int, hi: int, nseg: int, halo: int) -> None:
        """segs [n_local, rows, valid] = [halo slots | own block [lo, hi)].  Fill the ``n_halo`` halo
        slots with the segments just left of ``lo`` (owned by lower ranks) and send this rank's
        trailing segments to the ranks whose halo they are.  Every rank derives the same schedule."""
        if self.world == 1:
            return
        blocks = [self.block_of(r, nseg) for r in range(self.world)]

        def owner(g):
            return next(q for q, (ql, qh) in enumerate(blocks) if ql <= g < qh)

        ops = []
        for r, (l, h) in enumerate(blocks):
            if h == l:
                continue                                # empty block: owns no samples, needs no halo
            for g in range(max(0, l - halo), l):        # segment g is part of rank r's halo
                o = owner(g)
                if o == self.rank and r != self.rank:
                    ops.append(dist.P2POp(dist.isend, segs[n_halo + (g - lo)], self._peer(r), self.group))
                elif r == self.rank and o != self.rank:
                    ops.append(dist.P2POp(dist.irecv, segs[g - (lo - n_halo)], self._peer(o), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _peer(self, rank_in_group: int) -> int:
        return rank_in_group if self.group is None else dist.get_global_rank(self.group, rank_in_group)

    def combine(self, out: torch.Tensor, plan: tp.Optional[tp.Tuple[int, int, int]] = None) -> None:
        """Assemble the stems on every rank.  ``out`` [rows, L] holds this rank's sample ranges and zeros elsewhere.

        ``plan`` = (nseg, stride, length) when every pass used the same segment plan with no shift: rank q then
        owns exactly the samples [lo_q*stride, hi_q*stride) (the last rank up to ``length``), so the pieces are
        exchanged with ONE all-gather of compact [rows, max_len] buffers -- each rank receives (world-1)/world of
        the result once, half the traffic of the general path.  Otherwise (shift trick: the owned ranges move
        with the random offset of every pass) the zero-padded pieces are summed with an all-reduce.  Both give
        the bit pattern of a single-GPU run."""
        if self.world == 1:
            return
        if plan is None:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
            return
        nseg, stride, length = plan
        ranges = []
        for q in range(self.world):
            lo, hi = self.block_of(q, nseg)
            ranges.append((lo * stride, lo * stride) if hi == lo else
                          (lo * stride, length if hi >= nseg else hi * stride))
        width = max(b - a for a, b in ranges)
        rows = out.shape[0]
        a, b = ranges[self.rank]
        local = torch.zeros(rows, width, dtype=out.dtype, device=out.device)
        local[:, :b - a].copy_(out[:, a:b])
        gathered = torch.empty(self.world * rows, width, dtype=out.dtype, device=out.device)   # rank-major
        dist.all_gather_into_tensor(gathered, local, group=self.group)
        for q, (a, b) in enumerate(ranges):
            if q != self.rank and b > a:
                out[:, a:b].copy_(gathered[q * rows:(q + 1) * rows, :b - a])
