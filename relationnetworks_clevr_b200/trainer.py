"""Step-level host logic around the hot path (reference train.py:39-48, 256-258, 330).

* ``FlatClipAdam``: all parameters live in ONE flat fp32 buffer and all gradients in another; the backward of every op
  writes its parameter gradients straight into the flat gradient buffer (``ops.set_grad_sink``), so a step is
  [one NCCL all-reduce when torch.distributed is initialised] -> one fused clip_grad_norm + Adam(weight_decay) call
  (``ops.clip_adam_``, 2 launches).  Replaces the reference's ``clip_grad_norm`` + ``optim.Adam`` (~100 small launches)
  and ``nn.DataParallel``'s per-step broadcast / reduce.  The Adam step counter and the learning rate live on the device.
* ``train_step``: zero_grad / forward / nll_loss / backward / optimiser, returning the loss tensor.
* ``GraphedTrainStep``: the same step captured ONCE into a CUDA graph (static input buffers, device-side step counter,
  the gradient all-reduce inside the graph) and replayed: one host call per step instead of ~100 launches -- what the
  strong-scaling regime (80 questions per GPU, ~1.5 ms of device work) needs.
"""
from __future__ import annotations

import warnings
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
    """``sink=True`` (default): while this optimiser is the active one, parameter gradients are written by the ops'
    backward kernels directly into ``self.grad`` (``param.grad`` stays None).  Parameters whose gradient does not come
    from this library's ops (e.g. an nn.LSTM running in PyTorch) still arrive through ``param.grad`` and are copied in."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 5e-6, weight_decay: float = 1e-4,
                 clip_norm: float = 50.0, betas=(0.9, 0.999), eps: float = 1e-8, sink: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatClipAdam runs on CUDA parameters only")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)      # parameters become views of the flat buffer
            self.offsets.append(off)
            off += k
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.lr, self.weight_decay, self.clip_norm, self.betas, self.eps = lr, weight_decay, clip_norm, betas, eps
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)       # Adam step counter, incremented on the device
        self.lr_dev = torch.full((1,), lr, dtype=torch.float32, device=dev)
        self.norm_dev = torch.zeros(1, dtype=torch.float32, device=dev)     # pre-clip gradient norm of the last step
        self.last_norm: Optional[torch.Tensor] = None
        self.sink = sink
        if sink:
            ops.set_grad_sink(self.flat, self.grad)

    @property
    def step_count(self) -> int:
        return int(self.step_dev.item())

    def set_lr(self, lr: float) -> None:
        """Learning-rate schedule hook (train.py:332-340): updates the device-resident value the kernels read."""
        self.lr = lr
        self.lr_dev.fill_(lr)

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    def gather_grads(self) -> torch.Tensor:
        """Copy the gradients that arrived through autograd (``param.grad``) into the flat buffer.  With the sink active
        the library's ops have already written theirs in place and this touches only what is left; a parameter that
        received no gradient at all keeps its slice of the previous step unless the sink is off (then it is zeroed)."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is not None:
                self.grad[off:off + p.numel()].copy_(p.grad.reshape(-1))
            elif not self.sink:
                self.grad[off:off + p.numel()].zero_()
        return self.grad

    def check_aliasing(self) -> None:
        """A later ``model.cuda()`` / ``.float()`` / ``lstm.flatten_parameters()`` may re-allocate parameters and detach
        them from the flat buffer silently; the optimiser would then update memory the model no longer reads."""
        base, n = self.flat.data_ptr(), self.flat.numel() * 4
        for p, off in zip(self.params, self.offsets):
            if p.data_ptr() != base + 4 * off or not (base <= p.data_ptr() < base + n):
                raise RuntimeError("a parameter no longer aliases FlatClipAdam's flat buffer (was the model moved or "
                                   "flatten_parameters() called after the optimiser was built?)")

    def step(self) -> torch.Tensor:
        ops.join_aux(self.flat.device)         # gradients written on the auxiliary stream (question encoder) are complete
        g = self.gather_grads()
        scale = allreduce_flat_(g)
        self.last_norm = ops.clip_adam_(self.flat, g, self.exp_avg, self.exp_avg_sq, 0, self.lr, self.clip_norm,
                                        self.weight_decay, self.betas, self.eps, grad_scale=scale, step_dev=self.step_dev,
                                        lr_dev=self.lr_dev, total_out=self.norm_dev)
        return self.last_norm


def train_step(model, optimizer: FlatClipAdam, img, qst, label) -> torch.Tensor:
    """One iteration of the reference's training loop body (train.py:39-48)."""
    optimizer.zero_grad()
    output = model(img, qst)
    loss = F.nll_loss(output, label)
    loss.backward()
    optimizer.step()
    return loss


class GraphedTrainStep:
    """``train_step`` captured into a CUDA graph.

    ``step(img, qst, label)`` copies the batch into static device buffers (the copies are ordinary stream work: they may
    come from pinned host memory, non-blocking), replays the graph and returns the static loss tensor (valid until the
    next call).  Falls back to the eager ``train_step`` -- with a warning -- if the capture fails."""

    def __init__(self, model, optimizer: FlatClipAdam, img, qst, label, warmup: int = 3):
        self.model, self.opt = model, optimizer
        self.img, self.qst, self.label = img.clone(), qst.clone(), label.clone()
        self.graph = None
        self.loss = None
        optimizer.check_aliasing()
        # the warm-up steps below are real optimiser steps on the example batch: snapshot and restore everything they touch
        # (parameters, Adam moments, step counter, BatchNorm statistics) so building the graph does not train the model
        snap = [t.clone() for t in (optimizer.flat, optimizer.exp_avg, optimizer.exp_avg_sq, optimizer.step_dev)]
        bufs = [b for b in model.buffers()]
        snap_bufs = [b.clone() for b in bufs]
        side = torch.cuda.Stream(device=img.device)
        side.wait_stream(torch.cuda.current_stream(img.device))
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):                      # allocator / scratch / cudaFuncSetAttribute warm-up, off-graph
                    train_step(model, optimizer, self.img, self.qst, self.label)
            torch.cuda.current_stream(img.device).wait_stream(side)
            torch.cuda.synchronize(img.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.loss = train_step(model, optimizer, self.img, self.qst, self.label)
            self.graph = g
        except Exception as e:      # noqa: BLE001 -- any capture problem: stay correct, lose only the launch savings
            warnings.warn(f"CUDA-graph capture of the training step failed ({type(e).__name__}: {e}); running eagerly")
            self.graph = None
            torch.cuda.synchronize(img.device)
        with torch.no_grad():
            for dst, src in zip((optimizer.flat, optimizer.exp_avg, optimizer.exp_avg_sq, optimizer.step_dev), snap):
                dst.copy_(src)
            for dst, src in zip(bufs, snap_bufs):
                dst.copy_(src)

    @property
    def captured(self) -> bool:
        return self.graph is not None

    def step(self, img, qst, label) -> torch.Tensor:
        if self.graph is None:
            return train_step(self.model, self.opt, img, qst, label)
        if img.data_ptr() != self.img.data_ptr():
            self.img.copy_(img, non_blocking=True)
        if qst.data_ptr() != self.qst.data_ptr():
            self.qst.copy_(qst, non_blocking=True)
        if label.data_ptr() != self.label.data_ptr():
            self.label.copy_(label, non_blocking=True)
        self.graph.replay()
        return self.loss
