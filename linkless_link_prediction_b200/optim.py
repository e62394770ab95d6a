"""Flat-buffer Adam with per-group gradient clipping fused into the update — the optimiser tail of the
reference step (train_teacher_gnn.py:61-67,426-428; main.py:224-230): ``clip_grad_norm_(model, 1.0)``,
``clip_grad_norm_(predictor, 1.0)``, ``Adam.step()``.  Under data parallelism the one collective of a training
step, an NCCL all-reduce of the flat fp32 gradient bucket, happens here as well."""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional, Sequence

import torch

from . import _native as N


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.Adam(params, lr=...)`` (defaults betas=(0.9,0.999), eps=1e-8, no weight decay).

    Parameters are re-homed as views of one flat fp32 buffer (same for ``.grad``), so clipping + Adam is two
    launches regardless of the number of tensors and the DP gradient sync is one all-reduce."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None, distributed: bool = True):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._params: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        if not self._params or not all(p.is_cuda and p.dtype == torch.float32 for p in self._params):
            raise RuntimeError("FusedAdam needs fp32 CUDA parameters (no CPU fallback)")
        dev = self._params[0].device
        sizes = [p.numel() for p in self._params]
        self._offsets = [0]
        for s in sizes:
            self._offsets.append(self._offsets[-1] + s)
        n = self._offsets[-1]
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(self._params, self._offsets):
            view = self.flat_param[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self._step = 0
        self._dev_step = torch.zeros(1, dtype=torch.int64, device=dev)  # the step counter the kernel uses (graph safe)
        self.process_group = process_group
        self.distributed = distributed  # False: never all-reduce (single-rank reference runs inside a DP job)
        lib = N.load()
        self._ws = torch.empty(max(lib.llp_clip_adam_workspace_bytes(8), 256), dtype=torch.uint8, device=dev)
        self.group_norms = torch.zeros(8, dtype=torch.float32, device=dev)

    def reset_state(self) -> None:
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self._dev_step.zero_()
        self._step = 0

    def zero_grad(self, set_to_none: bool = False) -> None:  # grads stay views of the flat bucket
        self.flat_grad.zero_()
        for p, o in zip(self._params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def _group_bounds(self, clip_groups: Optional[Sequence[Iterable[torch.nn.Parameter]]]) -> List[int]:
        if not clip_groups:
            return [0, self._offsets[-1]]
        index = {id(p): i for i, p in enumerate(self._params)}
        bounds, expect = [0], 0
        for grp in clip_groups:
            ids = [index[id(p)] for p in grp]
            if ids != list(range(expect, expect + len(ids))):
                raise RuntimeError("clip groups must be contiguous runs of the optimizer's parameter order")
            expect += len(ids)
            bounds.append(self._offsets[expect])
        if expect != len(self._params):
            raise RuntimeError("clip groups must cover every parameter")
        return bounds

    @torch.no_grad()
    def step(self, closure=None, clip_groups=None, max_norm: float = 0.0):
        """``clip_groups``: lists of parameters clipped separately to ``max_norm`` before the update
        (the reference clips model and predictor separately)."""
        lib = N.require_gpu()
        grad_scale = 1.0
        if self.distributed and (self.process_group is not None or (
                torch.distributed.is_available() and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1)):
            import torch.distributed as dist
            world = dist.get_world_size(self.process_group)
            if world > 1:
                dist.all_reduce(self.flat_grad, group=self.process_group)  # NCCL sum; averaged inside the kernel
                grad_scale = 1.0 / world
        bounds = self._group_bounds(clip_groups)
        g = self.param_groups[0]
        self._step += 1
        arr = (ctypes.c_int64 * len(bounds))(*bounds)
        N.check(lib.llp_clip_adam(self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                  self.exp_avg_sq.data_ptr(), self.flat_param.numel(), arr, len(bounds) - 1, float(max_norm),
                                  grad_scale, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                  self._step, self._dev_step.data_ptr(), None, self.group_norms.data_ptr(), self._ws.data_ptr(),
                                  N.stream_ptr()),
                "llp_clip_adam")
        from . import ops
        if ops.compute_dtype() == torch.bfloat16:
            ops.prepare_weights(self._params)  # bf16 W and W^T of every weight matrix for the next step: one launch
        return None
