"""Training step and data parallelism for the MMBiDAF hot path.

The step is the reference's (train.py:146-155): forward -> loss.backward() -> clip_grad_norm_(max_grad_norm)
-> Adadelta(lr).step().  Data parallelism is one process per GPU (torchrun): every rank owns a shard of
the batch of videos -- the path has no cross-video coupling except the summed loss (models.py:170) and the
shared weights -- and the only collective is ONE all-reduce(SUM) of the flattened gradient per step over
NCCL / NVLink (the reference's nn.DataParallel, train.py:92, cannot actually scatter its list-typed lengths;
SURVEY.md section 2.1).  SUM, not mean: the reference loss is a sum over the batch, so summing rank gradients
reproduces the single-process gradient of the global batch.  The clip uses the global norm of the reduced
gradient, so every rank applies the identical update and replicas stay bit-identical.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .synth import Batch


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``n_items`` videos for ``rank`` (first ranks get the remainder)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_batch(batch: Batch, rank: int, world: int) -> Batch:
    """This rank's videos, re-padded to the shard's own maximum lengths (padding is masked / never visited)."""
    idx = list(shard_range(len(batch.text_len), rank, world))
    pick = lambda xs: [xs[i] for i in idx]
    tl, al, il, gl = pick(batch.text_len), pick(batch.audio_len), pick(batch.image_len), pick(batch.target_len)
    sel = torch.tensor(idx, dtype=torch.long)
    return Batch(batch.text[sel, :max(tl)].contiguous(), tl, batch.audio[sel, :max(al)].contiguous(), al,
                 batch.images[sel, :max(il)].contiguous(), il, batch.targets[sel, :max(gl)].contiguous(), gl, max(gl))


class FlatState:
    """All trainable parameters AND their gradients as views into two contiguous buffers: one all-reduce,
    one norm, one scale, and an optimiser that is five element-wise ops instead of 60 small-tensor updates."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat_param = torch.empty(total, dtype=first.dtype, device=first.device)
        self.flat = torch.zeros(total, dtype=first.dtype, device=first.device)        # gradients
        offset = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_param[offset:offset + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[offset:offset + n].view_as(p)
                p.grad = self.flat[offset:offset + n].view_as(p)
                offset += n

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce_sum(self, group=None) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    def clip_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ semantics (train.py:154) on the flat buffer; returns the norm."""
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm


FlatGrads = FlatState


class FlatAdadelta:
    """torch.optim.Adadelta (train.py:110: lr, rho 0.9, eps 1e-6, weight_decay) on the flat buffers."""

    def __init__(self, state: FlatState, lr: float = 1.0, rho: float = 0.9, eps: float = 1e-6, weight_decay: float = 0.0):
        self.s, self.lr, self.rho, self.eps, self.wd = state, lr, rho, eps, weight_decay
        self.square_avg = torch.zeros_like(state.flat)
        self.acc_delta = torch.zeros_like(state.flat)

    @torch.no_grad()
    def step(self) -> None:
        g = self.s.flat
        if self.wd != 0:
            g = g.add(self.s.flat_param, alpha=self.wd)
        self.square_avg.mul_(self.rho).addcmul_(g, g, value=1 - self.rho)
        delta = self.acc_delta.add(self.eps).sqrt_().div_(self.square_avg.add(self.eps).sqrt_()).mul_(g)
        self.acc_delta.mul_(self.rho).addcmul_(delta, delta, value=1 - self.rho)
        self.s.flat_param.add_(delta, alpha=-self.lr)


class Trainer:
    """One training step of an MMBiDAF-signature model; works single-process or under torch.distributed."""

    def __init__(self, model: torch.nn.Module, lr: float = 0.5, l2_wd: float = 0.0, max_grad_norm: float = 2.0,
                 group=None):
        self.model = model
        self.group = group
        self.max_grad_norm = max_grad_norm
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for p in model.parameters():                      # replicas start identical
                dist.broadcast(p.data, src=0, group=group)
        self.grads = FlatState(model.parameters())
        self.optimizer = FlatAdadelta(self.grads, lr=lr, weight_decay=l2_wd)                  # train.py:110
        self.last_grad_norm: Optional[torch.Tensor] = None

    def step(self, batch: Batch) -> torch.Tensor:
        """forward + backward + all-reduce + clip + Adadelta on this rank's shard; returns the local loss."""
        self.model.train()
        self.grads.zero()
        _, loss = self.model(batch.text, batch.text_len, batch.audio, batch.audio_len, batch.images, batch.image_len,
                             batch.targets, batch.target_len, batch.max_dec_len)
        loss.backward()
        self.grads.all_reduce_sum(self.group)
        self.last_grad_norm = self.grads.clip_(self.max_grad_norm)
        self.optimizer.step()
        return loss.detach()
