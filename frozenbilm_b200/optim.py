"""Fused clip_grad_norm_ + Adam over the flat trainable bucket (reference main.py:82-84: clip at
args.max_norm = 0.1, torch.optim.Adam with betas (0.9, 0.95), args.py:246-275), plus the data-parallel
gradient all-reduce the reference lacks: ONE NCCL all-reduce of the ~29.7 M-element bucket.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import lib


def allreduce_sum_and_scale(flat: torch.Tensor) -> float:
    """ONE collective over the flat gradient bucket (NCCL over NVLink on GPUs; gloo in the CPU tests).  Returns the
    1/world factor that the caller folds into the fused clip+Adam kernel instead of a separate scaling pass."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
        return 1.0 / dist.get_world_size()
    return 1.0


def broadcast_flat(flat: torch.Tensor, src: int = 0) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src)


def shard_seed(base_seed: int, rank: int, step_block: int) -> int:
    """Per-rank, per-batch seed of the synthetic data shard (each rank draws a different batch, weights share
    `base_seed`): the data-parallel partition of bench.py."""
    return base_seed + rank + 1000 * step_block


class FusedClipAdam:
    def __init__(self, model, lr=3e-5, betas=(0.9, 0.95), eps=1e-8, max_norm=0.1):
        self.model = model
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.step_count = 0
        self.m = self.v = None
        self.norm_sq = None

    def _state(self):
        p32, g32, p16 = self.model.flat_buffers()
        if self.m is None or self.m.numel() != p32.numel():
            self.m = torch.zeros_like(p32)
            self.v = torch.zeros_like(p32)
            self.norm_sq = torch.zeros(1, device=p32.device)
            self.norm_scratch = torch.zeros(1025, device=p32.device)   # partials + ticket of fb_sumsq_det
        return p32, g32, p16

    def zero_grad(self):
        _, g32, _ = self._state()
        g32.zero_()

    def step(self):
        """all-reduce (if a process group is up) -> global-norm clip -> Adam -> refresh bf16 operand copies."""
        p32, g32, p16 = self._state()
        # the module's backward already averaged the gradients over the ranks (overlapped with the backward itself);
        # only when that is switched off (model.dp_sync = False) is the bucket reduced here
        fl = self.model._flat
        if fl.get("synced"):
            scale = 1.0
            fl["synced"] = False
        else:
            scale = allreduce_sum_and_scale(g32)
        self.step_count += 1
        # fixed summation order: identical (all-reduced) gradients give a bit-identical clip factor on every rank
        lib.sumsq_det(g32, self.norm_sq, self.norm_scratch)
        lib.adam_step(p32, g32, self.m, self.v, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                      step=self.step_count, gnorm_sq=self.norm_sq, max_norm=self.max_norm, grad_scale=scale, p16=p16)
        self.model._mark_trainable_fresh()


def broadcast_trainable(model, src=0):
    """One-time broadcast of the trainable bucket: the reference seeds every rank differently BEFORE building the
    model (main.py:168-171), so unsynchronised replicas start from different adapters; this build deliberately
    starts all replicas from rank `src`'s values."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        p32, _, _ = model.flat_buffers()
        broadcast_flat(p32, src)
        # a c10d collective writes the tensor without bumping its autograd version counter: force the refresh of the
        # bf16 operand copies of the trainable tensors (ranks other than ``src`` would keep computing with their own)
        model._flat["version"] = -1
        model._ensure_engine()
