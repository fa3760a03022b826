"""Fused Adam: the optimizer of `ConvolutionalMulVAE` (reference model/disentangled_vae.py:304, stepped at
model/variational_base_vae.py:69) as ONE kernel launch over all parameter tensors (csrc/ops_optim.cu).

Drop-in for `torch.optim.Adam` on the options the reference uses (lr, default betas / eps, no weight decay, no amsgrad):
same constructor, same `state` layout (`step` CPU scalar tensor, `exp_avg`, `exp_avg_sq`), so `state_dict()` /
`load_state_dict()` interchange with torch's Adam and reference checkpoints.  CUDA parameters only: there is no CPU path.
"""
from __future__ import annotations

from typing import List

import torch

from . import lib

_CHUNK = 16384   # elements per block: 84 tensors / 61.4 M elements -> ~3800 blocks


class Adam(torch.optim.Adam):
    """A torch.optim.Adam (constructor, param_groups, state_dict) whose `step` is the fused kernel."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise NotImplementedError("dvae_b200.optim.Adam implements the reference's configuration: weight_decay=0, amsgrad=False")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        self._tables = {}     # per param group: device pointer / size / block tables
        self._grad_ptrs = {}  # per param group: the gradient addresses uploaded last
        # Non-finite gradients: the kernel leaves such elements untouched and raises a device flag, which is read back one
        # step late (pinned buffer + event: no host sync in the step).  `overflow_hook()` is then called -- the fp16 drop-in
        # model halves its gradient scale there -- and `overflow_steps` counts the occurrences.
        self.overflow_hook = None
        self.overflow_steps = 0
        self._inf_flag = None
        self._inf_host = None
        self._inf_event = None

    def _check_overflow(self, device):
        if self._inf_flag is None:
            self._inf_flag = torch.zeros(1, dtype=torch.int32, device=device)
            self._inf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._inf_event = torch.cuda.Event()
            return
        self._inf_event.synchronize()           # the previous step's flag: that step finished long ago
        if int(self._inf_host[0]) != 0:
            self.overflow_steps += 1
            if self.overflow_hook is not None:
                self.overflow_hook()
        self._inf_flag.zero_()

    def _build(self, gi: int, plist: List[torch.Tensor]):
        dev = plist[0].device
        for p in plist:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("dvae_b200.optim.Adam needs contiguous fp32 CUDA parameters (no CPU path)")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        sizes = [p.numel() for p in plist]
        blk_t, blk_o = [], []
        for t, n in enumerate(sizes):
            for off in range(0, n, _CHUNK):
                blk_t.append(t)
                blk_o.append(off)
        i64 = dict(dtype=torch.int64, device=dev)
        tb = dict(
            params=plist,
            p=torch.tensor([p.data_ptr() for p in plist], **i64),
            m=torch.tensor([self.state[p]["exp_avg"].data_ptr() for p in plist], **i64),
            v=torch.tensor([self.state[p]["exp_avg_sq"].data_ptr() for p in plist], **i64),
            g=torch.zeros(len(plist), **i64),
            sizes=torch.tensor(sizes, **i64),
            blk_t=torch.tensor(blk_t, dtype=torch.int32, device=dev),
            blk_o=torch.tensor(blk_o, **i64),
            nblk=len(blk_t),
            state_ptrs=[(self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr()) for p in plist],
        )
        self._tables[gi] = tb
        self._grad_ptrs[gi] = None
        return tb

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        launched = False
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            tb = self._tables.get(gi)
            if tb is None or len(tb["params"]) != len(plist) or any(a is not b for a, b in zip(tb["params"], plist)) or \
                    any((self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr()) != sp
                        for p, sp in zip(plist, tb["state_ptrs"])):     # e.g. after load_state_dict
                tb = self._build(gi, plist)
            gptrs = []
            for p in plist:
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device or g.is_sparse:
                    raise RuntimeError("dvae_b200.optim.Adam needs dense contiguous fp32 gradients on the parameter's device")
                gptrs.append(g.data_ptr())
            if gptrs != self._grad_ptrs[gi]:     # the allocator hands back the same blocks step after step: rare upload
                tb["g"].copy_(torch.tensor(gptrs, dtype=torch.int64))
                self._grad_ptrs[gi] = gptrs
            if not launched:
                self._check_overflow(plist[0].device)       # once per step, before the first launch
                launched = True
            st0 = self.state[plist[0]]["step"]
            step = int(st0.item()) + 1
            beta1, beta2 = group["betas"]
            lib.call("dvae_adam_step", lib.ptr(tb["p"]), lib.ptr(tb["g"]), lib.ptr(tb["m"]), lib.ptr(tb["v"]), lib.ptr(tb["sizes"]),
                     lib.ptr(tb["blk_t"]), lib.ptr(tb["blk_o"]), tb["nblk"], _CHUNK, float(group["lr"]), float(beta1), float(beta2),
                     float(group["eps"]), step, lib.ptr(self._inf_flag), lib.stream())
            # the kernel wrote the parameters through raw pointers: tell autograd (and every cache keyed on `_version`,
            # e.g. the tensor-core weight copies of the drop-in modules) that they changed
            torch.autograd.graph.increment_version(plist)
            for p in plist:
                self.state[p]["step"] += 1     # CPU scalars, like torch's non-capturable Adam
        if launched:
            self._inf_host.copy_(self._inf_flag, non_blocking=True)
            self._inf_event.record()
        return loss
