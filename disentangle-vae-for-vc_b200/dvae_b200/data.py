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
    """Batch sampler for data-parallel training: every step draws `world * pairs_per_rank / group_size` DISTINCT speakers
    with `group_size` rows each and hands every rank the rows of `pairs_per_rank / group_size` whole speakers, sorted by
    speaker -- so each rank gets EXACTLY `pairs_per_rank` rows every step (equal shards: the all-reduce average of the
    per-rank losses is the global-batch loss, and shapes never change), no speaker is split across ranks, and the
    speaker-group kernels never need a cross-rank exchange.

    Per epoch every speaker's items are shuffled and cut into chunks of `group_size` (a remainder shorter than that is
    dropped for the epoch); the chunks are shuffled and dealt into steps greedily so that a step never holds two chunks of
    one speaker; chunks that cannot complete a step are dropped.  All ranks must construct the sampler with the same
    `speaker_ids` / `seed` and call `set_epoch` with the same epoch: the plan is then identical everywhere.
    `group_size` defaults to the largest divisor of `pairs_per_rank` that every speaker can fill (BASELINE configs:
    8 utterances per speaker).  Use as `DataLoader(ds, batch_sampler=...)`.  The reference (`train.py:49-58`) uses a plain
    shuffled DataLoader on one process."""

    def __init__(self, speaker_ids, pairs_per_rank: int, rank: int = 0, world: int = 1, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = True, group_size: Optional[int] = None):
        import numpy as np
        self.ids = np.asarray(speaker_ids).reshape(-1)
        self.pairs_per_rank, self.rank, self.world = int(pairs_per_rank), int(rank), int(world)
        self.shuffle, self.seed, self.drop_last, self.epoch = shuffle, seed, drop_last, 0
        if not 0 <= self.rank < self.world:
            raise ValueError("rank must be in [0, world)")
        if not drop_last:
            raise ValueError("equal shards need drop_last=True (a partial step cannot give every rank pairs_per_rank rows)")
        speakers, counts = np.unique(self.ids, return_counts=True)
        if group_size is None:
            smallest = int(counts.min())
            group_size = max(d for d in range(1, self.pairs_per_rank + 1) if self.pairs_per_rank % d == 0 and d <= smallest)
        self.group_size = int(group_size)
        if self.group_size < 1 or self.pairs_per_rank % self.group_size != 0:
            raise ValueError("group_size must divide pairs_per_rank")
        self.groups_per_step = self.world * self.pairs_per_rank // self.group_size
        usable = int((counts >= self.group_size).sum())
        if usable < self.groups_per_step:
            raise ValueError(f"a step needs {self.groups_per_step} distinct speakers with >= {self.group_size} items each, "
                             f"the dataset has {usable}: no rank may receive an empty shard")
        self._plan_cache = (None, None)

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def _plan(self):
        """Steps of this epoch: a list of [groups_per_step][group_size] index arrays, speakers sorted within a step."""
        import numpy as np
        if self._plan_cache[0] == self.epoch:
            return self._plan_cache[1]
        rng = np.random.default_rng(self.seed + self.epoch)
        chunks = []                                    # (speaker, indices)
        for spk in np.unique(self.ids):
            idx = np.flatnonzero(self.ids == spk)
            if self.shuffle:
                idx = idx[rng.permutation(len(idx))]
            for c in range(len(idx) // self.group_size):
                chunks.append((spk, idx[c * self.group_size:(c + 1) * self.group_size]))
        order = rng.permutation(len(chunks)) if self.shuffle else np.arange(len(chunks))
        pending = [chunks[i] for i in order]
        steps = []
        while True:
            taken, used, rest = [], set(), []
            for ch in pending:
                if len(taken) < self.groups_per_step and ch[0] not in used:
                    taken.append(ch)
                    used.add(ch[0])
                else:
                    rest.append(ch)
            if len(taken) < self.groups_per_step:
                break
            taken.sort(key=lambda ch: ch[0])           # rows of a step sorted by speaker: contiguous group ranges per rank
            steps.append(taken)
            pending = rest
        self._plan_cache = (self.epoch, steps)
        return steps

    def __len__(self) -> int:
        return len(self._plan())

    def __iter__(self):
        k = self.pairs_per_rank // self.group_size     # whole speakers per rank and step
        for step in self._plan():
            mine = step[self.rank * k:(self.rank + 1) * k]
            yield [int(i) for _, idx in mine for i in idx]
