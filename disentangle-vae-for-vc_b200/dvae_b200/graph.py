"""One training step (weight-copy refresh + forward + loss + backward) as ONE CUDA-graph launch.

The step is ~550 kernel launches from Python / ctypes: ~6 ms of host time per step next to ~18 ms of GPU time.  On one GPU the
host is not the floor, but it is the part of the step that is exposed to the rest of the machine (data loading, other ranks'
jitter).  The hand-scheduled engine already is a static launch sequence for a given batch shape -- streams, events, memset
nodes and kernels with raw-pointer arguments -- so it can be captured as it is (`torch.cuda.graph`: allocations inside the
capture come from the graph's private pool and keep their addresses; the side stream of the engine forks from and joins the
capturing stream).  What stays OUTSIDE the graph:

  * the optimizer step: Adam's bias corrections are host scalars passed by value (`dvae_b200.optim.Adam`, one launch), and
    it bumps the parameters' versions; the graph therefore begins with an unconditional refresh of the tensor-core weight copies;
  * the reparameterisation noise: drawn on the CPU in the reference's order (or by the model's `noise_hook`) and copied into
    static device tensors before every replay;
  * the inputs: copied into static device tensors.

Reference call sequence this replaces: `VariationalBaseModelVAE.step` (model/variational_base_vae.py:58-70) minus
`optimizer.step()`.
"""
from __future__ import annotations

from typing import List

import torch


class GraphedTrainStep:
    RING = 4

    def __init__(self, wrapper, x1: torch.Tensor, x2: torch.Tensor, warmup: int = 1):
        model = wrapper.model
        if not model.training:
            raise RuntimeError("GraphedTrainStep captures the training step: call model.train() first")
        if getattr(model._engine, "buckets", None) is not None:
            raise RuntimeError("GraphedTrainStep: gradient buckets (data-parallel all-reduce) are not captured; use the eager step")
        self.wrapper, self.model = wrapper, model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.shape = (tuple(x1.shape), tuple(x2.shape))
        self.x1 = x1.detach().to(torch.float32).clone()
        self.x2 = x2.detach().to(torch.float32).clone()
        user_hook = model.noise_hook
        shapes: List[tuple] = []

        def learn(shape):
            shapes.append(tuple(shape))
            return torch.randn(tuple(shape), device=self.x1.device)
        # eager warm-up on a side stream (as torch.cuda.graph requires): the library's one-time set-up calls (function attributes,
        # occupancy queries, symbol addresses), the learned gradient-buffer layout and the noise shapes of one step
        # (the warm-up steps must not count as training steps: BatchNorm's running statistics are put back afterwards)
        s = torch.cuda.Stream(device=self.x1.device)
        s.wait_stream(torch.cuda.current_stream())
        model.noise_hook = learn
        try:
            with torch.cuda.stream(s):
                saved = [(b, b.detach().clone()) for b in model.buffers()]
                for _ in range(max(1, warmup)):
                    shapes.clear()
                    self._clear_grads()
                    self._eager(self.x1, self.x2)
                with torch.no_grad():
                    for b, keep in saved:
                        b.copy_(keep)
            torch.cuda.current_stream().wait_stream(s)
            self.noise = [torch.empty(sh, device=self.x1.device, dtype=torch.float32) for sh in shapes]
            # host staging of the CPU-drawn noise: a ring, because the host may run several replays ahead of the GPU
            self._pinned = [[torch.empty(sh, dtype=torch.float32, pin_memory=True) for sh in shapes] for _ in range(self.RING)]
            self._pin_events = [None] * self.RING
            self._calls = 0
            k = [0]

            def static(shape):
                t = self.noise[k[0] % len(self.noise)]
                k[0] += 1
                assert tuple(t.shape) == tuple(shape), "noise draws changed between warm-up and capture"
                return t
            model.noise_hook = static
            # the refresh of the tensor-core weight copies must be IN the graph (it runs when a parameter version has moved)
            torch.autograd.graph.increment_version(self.params)
            self._clear_grads()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs, losses = self._eager(self.x1, self.x2)
                self.losses = torch.stack([l.detach() for l in losses])
        finally:
            model.noise_hook = user_hook
        self.grads = [p.grad for p in self.params]

    def _clear_grads(self) -> None:
        for p in self.params:
            p.grad = None

    def _eager(self, x1, x2):
        out = self.model(x1, x2)
        losses = self.wrapper.loss_functionGVAE2(x1, x2, *out, train=True)
        losses[0].backward()
        return out, losses

    def __call__(self, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
        """Runs the captured step on (x1, x2).  Returns the 8 loss terms as ONE device tensor (valid until the next call); the
        parameter gradients are in `p.grad` (overwritten, not accumulated); the model outputs of the step are `self.outputs`."""
        if (tuple(x1.shape), tuple(x2.shape)) != self.shape:
            raise ValueError(f"GraphedTrainStep was captured for shapes {self.shape}, got {tuple(x1.shape)}, {tuple(x2.shape)}")
        self.x1.copy_(x1, non_blocking=True)
        self.x2.copy_(x2, non_blocking=True)
        hook = self.model.noise_hook
        slot = self._calls % self.RING
        self._calls += 1
        if hook is None and self._pin_events[slot] is not None:
            self._pin_events[slot].synchronize()   # the copies that read this staging set (RING calls ago) have run
        for dst, pin in zip(self.noise, self._pinned[slot]):
            if hook is not None:
                dst.copy_(hook(tuple(dst.shape)), non_blocking=True)
            else:
                pin.normal_()                      # CPU default generator, in the reference's draw order
                dst.copy_(pin, non_blocking=True)
        if hook is None:
            if self._pin_events[slot] is None:
                self._pin_events[slot] = torch.cuda.Event()
            self._pin_events[slot].record()
        for p, g in zip(self.params, self.grads):
            if p.grad is not g:
                p.grad = g
        self.graph.replay()
        return self.losses
