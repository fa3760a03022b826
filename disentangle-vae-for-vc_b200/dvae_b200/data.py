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
