"""Host -> device input pipeline for the training loop.

The reference's loop (`model/variational_base_vae.py:74-101`) moves every batch with a blocking `.to(device)` right
before the step, so the 21 MB of a BASELINE config-2 batch (~1 ms over PCIe) sit on the critical path of every step.
`DevicePrefetcher` wraps any iterable of `(mel1, mel2, speaker_ids)` batches: batch k+1 is copied on a side stream
(from pinned memory, `non_blocking`) while batch k is being computed; consuming a batch makes the compute stream wait
for that batch's copy event only.  `AsyncScalars` is the matching device -> host side: scalars of step k are copied to
pinned memory without blocking and read one step later, when they have long arrived (one 32-byte copy per step instead
of eight `.item()` synchronisations, `model/variational_base_vae.py:70`).
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import torch


class DevicePrefetcher:
    def __init__(self, loader: Iterable, device: torch.device, depth: int = 2, pin: bool = True):
        self.loader, self.device, self.depth, self.pin = loader, torch.device(device), max(1, depth), pin
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches for a CUDA device (dvae_b200 has no CPU path)")
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        mel1, mel2, *rest = batch
        out = []
        with torch.cuda.stream(self.stream):
            for t in (mel1, mel2):
                if self.pin and not t.is_pinned():
                    t = t.pin_memory()
                d = t.to(self.device, non_blocking=True)
                out.append(d if d.dtype == torch.float32 else d.float())
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return (out[0], out[1], *rest), ev

    def __iter__(self) -> Iterator[Tuple]:
        it = iter(self.loader)
        queue = []
        try:
            while len(queue) < self.depth:
                queue.append(self._stage(next(it)))
        except StopIteration:
            it = None
        while queue:
            batch, ev = queue.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in batch[:2]:
                t.record_stream(cur)      # allocated on the copy stream, consumed on the compute stream
            if it is not None:
                try:
                    queue.append(self._stage(next(it)))
                except StopIteration:
                    it = None
            yield batch


class AsyncScalars:
    """Non-blocking read-back of a small device tensor per step; `push` returns the values of the PREVIOUS push."""

    def __init__(self, numel: int, device: torch.device, slots: int = 2):
        self.bufs = [torch.empty(numel, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.events = [torch.cuda.Event() for _ in range(slots)]
        self.k = 0
        self.pending: Optional[int] = None

    def push(self, values: torch.Tensor) -> Optional[list]:
        slot = self.k % len(self.bufs)
        self.bufs[slot].copy_(values.detach().reshape(-1), non_blocking=True)
        self.events[slot].record()
        prev, self.pending = self.pending, slot
        self.k += 1
        return self._read(prev)

    def flush(self) -> Optional[list]:
        prev, self.pending = self.pending, None
        return self._read(prev)

    def _read(self, slot):
        if slot is None:
            return None
        self.events[slot].synchronize()
        return self.bufs[slot].tolist()


class SpeakerGroupBatchSampler(torch.utils.data.Sampler):
    """Batch sampler for data-parallel training: every step draws a global batch of `world * pairs_per_rank` items and
    hands each rank the rows of WHOLE speaker groups (`parallel.shard_pairs_by_speaker`), sorted by speaker, so that the
    speaker-group kernels never need a cross-rank exchange.  All ranks must construct it with the same
    `speaker_ids`, `seed` and call `set_epoch` with the same epoch: the global permutation is then identical everywhere.

    With equally sized speaker groups (the BASELINE configs: 8 utterances per speaker) every rank gets exactly
    `pairs_per_rank` rows; otherwise ranks differ by at most one group.  Use as `DataLoader(ds, batch_sampler=...)`.
    The reference (`train.py:49-58`) uses a plain shuffled DataLoader on one process."""

    def __init__(self, speaker_ids, pairs_per_rank: int, rank: int = 0, world: int = 1, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = True):
        import numpy as np
        self.ids = np.asarray(speaker_ids).reshape(-1)
        self.pairs_per_rank, self.rank, self.world = int(pairs_per_rank), int(rank), int(world)
        self.shuffle, self.seed, self.drop_last, self.epoch = shuffle, seed, drop_last, 0
        if not 0 <= self.rank < self.world:
            raise ValueError("rank must be in [0, world)")

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def __len__(self) -> int:
        g = self.pairs_per_rank * self.world
        return len(self.ids) // g if self.drop_last else -(-len(self.ids) // g)

    def __iter__(self):
        import numpy as np
        from .parallel import shard_pairs_by_speaker
        n, g = len(self.ids), self.pairs_per_rank * self.world
        order = np.random.default_rng(self.seed + self.epoch).permutation(n) if self.shuffle else np.arange(n)
        for b in range(len(self)):
            glob = order[b * g:(b + 1) * g]
            local = shard_pairs_by_speaker(self.ids[glob], self.rank, self.world)
            yield glob[local].tolist()
