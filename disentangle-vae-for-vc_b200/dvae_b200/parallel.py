"""Data-parallel layer: whole speaker groups per rank, bucketed gradient all-reduce overlapped with backward.

The reference is single-process; this is the new part named by north_star (3).  One process per GPU,
`torch.distributed` (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests) is the plumbing:

  * `shard_pairs_by_speaker`  -- host logic: rows (pairs) are bucketed by speaker id and WHOLE speaker groups are
    assigned to ranks as contiguous group ranges balanced by row count, so the speaker-group kernels never need a
    cross-rank exchange.
  * `GradBuckets`             -- the engine writes every parameter gradient straight into one of a few flat fp32
    buckets laid out in backward order (postnet -> decoder -> encoder).  As soon as the last gradient of a bucket
    has been produced, an all-reduce (average) of that bucket is enqueued on a side stream, so communication of
    bucket k overlaps the backward compute of bucket k+1.  `finish()` joins the side stream.

The path has no other collective: BatchNorm statistics stay per rank (the reference has no SyncBN).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

# gradient production order of engine.backward, coarsely: bucket boundaries follow the module groups
BUCKET_PREFIXES: List[Tuple[str, ...]] = [
    ("postnet.",),
    ("dec_linear2.", "dec_lstm2."),
    ("dec_modules.", "dec_lstm1."),
    ("dec_pre_linear2.", "dec_pre_linear1."),
    ("style.", "content.", "enc_linear."),
    ("enc_lstm.", "enc_modules."),
]


def group_counts(speaker_ids: Sequence[int]):
    """(order-of-first-occurrence group index per row, rows per group) -- integer bookkeeping, bit exact."""
    order: Dict[int, int] = {}
    gid = np.empty(len(speaker_ids), dtype=np.int64)
    for i, s in enumerate(np.asarray(speaker_ids).reshape(-1).tolist()):
        gid[i] = order.setdefault(s, len(order))
    return gid, np.bincount(gid, minlength=len(order)).astype(np.int64)


def assign_groups_to_ranks(counts: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous group ranges [first, last) per rank, balanced by cumulative row count; every rank gets at least one
    group when there are at least `world` groups."""
    counts = [int(c) for c in counts]
    total, G = sum(counts), len(counts)
    out, g, acc = [], 0, 0
    for r in range(world):
        start = g
        target = total * (r + 1) / world
        while g < G and (G - g) > (world - 1 - r) and (g == start or acc + counts[g] <= target + 1e-9):
            acc += counts[g]
            g += 1
        if r == world - 1:
            g = G
        out.append((start, g))
    return out


def shard_pairs_by_speaker(speaker_ids: Sequence[int], rank: int, world: int) -> np.ndarray:
    """Row indices (into the global batch) owned by `rank`: rows sorted by speaker group (stable), whole groups per
    rank.  The union over ranks is a permutation of all rows."""
    gid, counts = group_counts(speaker_ids)
    order = np.argsort(gid, kind="stable")
    ranges = assign_groups_to_ranks(counts, world)
    offs = np.concatenate([[0], np.cumsum(counts)])
    a, b = ranges[rank]
    return order[offs[a]:offs[b]]


class GradBuckets:
    """Flat fp32 gradient buckets + overlapped all-reduce.  Works with any backend (CPU tensors + gloo in tests)."""

    def __init__(self, named_shapes: Sequence[Tuple[str, Tuple[int, ...]]], device, process_group=None,
                 comm_stream: Optional["torch.cuda.Stream"] = None):
        self.pg = process_group
        self.device = torch.device(device)
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.slots: Dict[str, Tuple[int, int, Tuple[int, ...]]] = {}
        sizes = [0] * len(BUCKET_PREFIXES)
        for name, shape in named_shapes:
            b = self.bucket_of(name)
            n = int(np.prod(shape)) if len(shape) else 1
            n_pad = (n + 63) // 64 * 64            # keep every slot 256-byte aligned (vector loads / TMA-friendly)
            self.slots[name] = (b, sizes[b], tuple(shape))
            sizes[b] += n_pad
        self.sizes = sizes
        self.members: List[List[str]] = [[n for n, s in self.slots.items() if s[0] == b] for b in range(len(sizes))]
        self.flat: List[torch.Tensor] = [torch.zeros(max(s, 1), device=self.device, dtype=torch.float32) for s in sizes]
        self.use_cuda = self.device.type == "cuda"
        self.comm_stream = comm_stream if comm_stream is not None else (torch.cuda.Stream(self.device) if self.use_cuda else None)
        self._pending: List[set] = []
        self._works = []
        self.extra_streams: List["torch.cuda.Stream"] = []   # other streams that also produce gradients (engine side stream)
        self.begin()

    @staticmethod
    def bucket_of(name: str) -> int:
        for b, prefixes in enumerate(BUCKET_PREFIXES):
            if name.startswith(prefixes):
                return b
        raise KeyError(f"parameter {name!r} does not belong to any gradient bucket")

    def begin(self) -> None:
        """Start a backward pass: zero the buckets (split-K epilogues accumulate into them)."""
        for f in self.flat:
            f.zero_()
        self._pending = [set(m) for m in self.members]
        self._works = []

    def view(self, name: str) -> torch.Tensor:
        b, off, shape = self.slots[name]
        n = int(np.prod(shape)) if len(shape) else 1
        return self.flat[b][off:off + n].view(shape)

    def ready(self, name: str) -> None:
        """The gradient of `name` is final.  Launches the bucket's all-reduce when it was the last one missing."""
        b = self.slots[name][0]
        self._pending[b].discard(name)
        if not self._pending[b] and self.world > 1:
            self._launch(b)

    def _launch(self, b: int) -> None:
        buf = self.flat[b]
        if self.use_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            for st in self.extra_streams:
                self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._works.append((w, buf))

    def finish(self) -> None:
        """Join communication with the compute stream (call once at the end of backward)."""
        if self.world <= 1:
            return
        missing = [b for b, p in enumerate(self._pending) if p]
        for b in missing:       # defensive: a bucket whose gradients were not all reported is still reduced
            self._pending[b] = set()
            self._launch(b)
        if self.use_cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        else:
            for w, buf in self._works:
                w.wait()
                buf.div_(self.world)
            self._works = []

    def payload_bytes(self) -> int:
        return 4 * sum(self.sizes)
