"""Step-level host logic around the hot path (reference train.py:39-48, 256-258, 330).

* ``FlatClipAdam``: all parameters live in ONE flat fp32 buffer; a step is (gather grads into a flat
  buffer) -> [one NCCL all-reduce when torch.distributed is initialised] -> one fused
  clip_grad_norm + Adam(weight_decay) kernel (``ops.clip_adam_``).  Replaces the reference's
  ``clip_grad_norm`` + ``optim.Adam`` (~100 small launches) and ``nn.DataParallel``'s per-step
  broadcast/reduce.
* ``train_step``: zero_grad / forward / nll_loss / backward / optimiser, returning the loss tensor.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops


def shard_rows(n: int, rank: int, world: int) -> slice:
    """Rows of a global batch owned by `rank`: contiguous equal shards, rank r takes [r*n/W, (r+1)*n/W)
    (the DataParallel-equivalent split of train.py:256-258)."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return slice(rank * per, (rank + 1) * per)


def allreduce_flat_(flat: torch.Tensor) -> float:
    """The single exchange step of the path: sum-all-reduce of the flat gradient buffer (NCCL on GPUs, gloo in
    the CPU tests).  Returns the scale (1/world) the optimiser applies so the result is the mean over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


class FlatClipAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 5e-6, weight_decay: float = 1e-4,
                 clip_norm: float = 50.0, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatClipAdam runs on CUDA parameters only")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)      # parameters become views of the flat buffer
            off += k
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.lr, self.weight_decay, self.clip_norm, self.betas, self.eps = lr, weight_decay, clip_norm, betas, eps
        self.step_count = 0
        self.last_norm: Optional[torch.Tensor] = None

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    def gather_grads(self) -> torch.Tensor:
        """Flatten .grad of every parameter into the flat gradient buffer (zeros where unused)."""
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        torch.cat(pieces, out=self.grad)
        return self.grad

    def step(self) -> torch.Tensor:
        g = self.gather_grads()
        scale = allreduce_flat_(g)
        self.step_count += 1
        self.last_norm = ops.clip_adam_(self.flat, g, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                                        self.clip_norm, self.weight_decay, self.betas, self.eps, grad_scale=scale)
        return self.last_norm


def train_step(model, optimizer: FlatClipAdam, img, qst, label) -> torch.Tensor:
    """One iteration of the reference's training loop body (train.py:39-48)."""
    optimizer.zero_grad()
    output = model(img, qst)
    loss = F.nll_loss(output, label)
    loss.backward()
    optimizer.step()
    return loss
