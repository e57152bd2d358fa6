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

from .functional import leaf_lanes
from .synth import Batch


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``n_items`` videos for ``rank`` (first ranks get the remainder)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_batch(batch: Batch, rank: int, world: int) -> Batch:
    """This rank's videos.  Audio and key-frames are re-padded to the shard's own maxima: their padding is masked in both
    BiDAF soft-maxes (attention.py:43-44) and never visited by the recurrences.  Text and targets keep the GLOBAL batch's
    padded widths: the decoder's two attention soft-maxes and its coverage term run over every padded text position
    (attention.py:148,154, quirk Q2), and the loss is averaged over the padded number of decode steps with padded steps scoring
    sentence 0 (models.py:168,179, quirk Q4) -- a narrower shard would change alpha, the loss and every gradient, and the summed
    rank gradients would no longer be the single-process gradient of the global batch."""
    idx = list(shard_range(len(batch.text_len), rank, world))
    pick = lambda xs: [xs[i] for i in idx]
    tl, al, il, gl = pick(batch.text_len), pick(batch.audio_len), pick(batch.image_len), pick(batch.target_len)
    sel = torch.tensor(idx, dtype=torch.long)
    return Batch(batch.text[sel].contiguous(), tl, batch.audio[sel, :max(al)].contiguous(), al,
                 batch.images[sel, :max(il)].contiguous(), il, batch.targets[sel].contiguous(), gl, batch.max_dec_len)


def _lstm_directions_adjacent(params: List[torch.nn.Parameter]) -> List[torch.nn.Parameter]:
    """nn.LSTM registers a bidirectional layer as (w_ih, w_hh, b_ih, b_hh, w_ih_reverse, w_hh_reverse, b_ih_reverse, b_hh_reverse).
    Re-ordered here to (w_ih, w_ih_reverse, w_hh, w_hh_reverse, ...): in the flat buffer the two directions of every tensor are then
    contiguous, and the recurrence's stacked operands are views instead of per-step concatenations (functional._stacked).  Groups are
    recognised by shape: eight consecutive parameters with shapes (a, b, c, c, a, b, c, c), c one-dimensional, every size a multiple of
    four (no padding between them).  Anything else keeps its place; the order is the same on every rank."""
    out, i = [], 0
    while i < len(params):
        g = params[i:i + 8]
        if (len(g) == 8 and g[0].dim() == 2 and g[1].dim() == 2 and g[2].dim() == 1 and g[2].shape == g[3].shape
                and all(g[k].shape == g[k + 4].shape for k in range(4)) and g[0].shape[0] == g[2].shape[0] == g[1].shape[0]
                and all(t.numel() % 4 == 0 for t in g)):
            out += [g[0], g[4], g[1], g[5], g[2], g[6], g[3], g[7]]
            i += 8
        elif (len(g) == 8 and all(g[k].dim() == 2 and g[k].shape == g[0].shape and g[k].shape[0] == g[k].shape[1] for k in (0, 2, 4, 6))
              and all(g[k].dim() == 1 and g[k].shape[0] == g[0].shape[0] for k in (1, 3, 5, 7)) and g[0].shape[0] % 4 == 0):
            # a two-layer HighwayEncoder: transforms (W, b) x 2 then gates (W, b) x 2 (layers/encoding.py) -> per layer gate and
            # transform side by side, weights then biases: the stacked (2H, H) operand of a layer's one GEMM is a view
            out += [g[4], g[0], g[5], g[1], g[6], g[2], g[7], g[3]]
            i += 8
        else:
            out.append(params[i])
            i += 1
    return out


class FlatState:
    """All trainable parameters AND their gradients as views into two contiguous buffers: one all-reduce,
    one norm, one scale, and an optimiser that is five element-wise ops instead of 60 small-tensor updates."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.params = _lstm_directions_adjacent(self.params)
        pad = lambda n: (n + 3) // 4 * 4                 # every parameter stays 16-byte aligned (kernels use float4)
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += pad(p.numel())
        first = self.params[0]
        self.flat_param = torch.zeros(total, dtype=first.dtype, device=first.device)
        self.flat = torch.zeros(total, dtype=first.dtype, device=first.device)        # gradients
        with torch.no_grad():
            for p, offset in zip(self.params, self.offsets):
                n = p.numel()
                self.flat_param[offset:offset + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[offset:offset + n].view_as(p)
                p.grad = self.flat[offset:offset + n].view_as(p)

    def zero(self) -> None:
        self.flat.zero_()

    def pack(self, grads) -> None:
        """Copy a tuple of per-parameter gradients (None = zero) into the flat buffer with one launch of our own kernel
        (``mmb_pack_segments``: the source pointers travel as kernel parameters; 12.8 MB).  On the CPU (gloo tests) and for
        non-fp32 buffers: one batched concatenation."""
        if self.flat.is_cuda and self.flat.dtype == torch.float32:
            import ctypes

            from . import _lib, ops
            n = len(self.params)
            keep = [None if g is None else (g if g.is_contiguous() else g.contiguous()) for g in grads]
            for g in keep:
                assert g is None or g.dtype == torch.float32
            srcs = (ctypes.c_void_p * n)(*[None if g is None else g.data_ptr() for g in keep])
            if not hasattr(self, "_pack_tables"):
                self._pack_tables = ((ctypes.c_longlong * n)(*self.offsets), (ctypes.c_longlong * n)(*[p.numel() for p in self.params]))
            offs, sizes = self._pack_tables
            _lib.check(_lib.lib().mmb_pack_segments(srcs, offs, sizes, n, _lib.ptr(self.flat), _lib.stream()), "mmb_pack_segments")
            ops._count(1)
            return
        pieces = []
        for g, p in zip(grads, self.params):
            n = p.numel()
            pieces.append(g.reshape(-1) if g is not None else p.new_zeros(n))
            if n % 4:
                pieces.append(p.new_zeros(4 - n % 4))
        torch.cat(pieces, out=self.flat)

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

    # ---- CUDA-graph replay of the whole step (launch-bound at ~1200 kernels per step) ----------------------
    def capture(self, batch: Batch, warmup: int = 3, error_mode: str = "global") -> None:
        """Capture forward + backward + all-reduce + clip + Adadelta for ``batch`` (resident on the device) into
        one CUDA graph.  Later ``step_graphed(other)`` copies a batch of the SAME padded shapes and lengths into
        the captured input tensors and replays.  Dropout draws fresh masks on every replay (torch's generator
        is graph-aware)."""
        import gc
        from .layers.encoding import pin_lengths
        self._static = batch
        # static length tensors (kept alive here): the captured kernels read lengths, masks and orders from fixed addresses that
        # step_graphed() rewrites for every new batch
        self._plans = [pin_lengths(lst, batch.text.device) for lst in (batch.text_len, batch.audio_len, batch.image_len)]
        # drop per-sequence caches that still own tensors (and stream-usage records) from eager steps, so that
        # nothing created on another stream is released while the capture is in flight
        for m in self.model.modules():
            if hasattr(m, "_cache"):
                m._cache = None
        gc.collect()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(batch)
        torch.cuda.current_stream().wait_stream(side)
        # The NCCL all-reduce stays outside the graphs (one eager call on the static flat buffer between two
        # replays): graph A = forward + backward + pack, graph B = clip + Adadelta.
        # The step is captured on a stream that ranks ABOVE the weight-gradient lanes (functional.leaf_lanes: default priority) and below
        # the model's encoder chains (models.py): kernel nodes keep their stream's priority, so whenever a chain kernel and a leaf GEMM
        # are both ready the chain goes first (bench.py, config 3: 4.40 -> 4.34 ms per step with the chains above the leaves).
        import os
        self._graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(priority=int(os.environ.get("MMB_MAIN_PRIO", "-1")))
        with torch.cuda.graph(self._graph, stream=cap, capture_error_mode=error_mode):
            self._static_loss = self._forward_backward(batch)
        self._graph_update = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_update, capture_error_mode=error_mode):
            self._update()

    def step_graphed(self, batch: Optional[Batch] = None) -> torch.Tensor:
        if batch is not None and batch is not self._static:
            s = self._static
            same_bucket = all(a.shape == b.shape for a, b in ((batch.text, s.text), (batch.audio, s.audio),
                                                              (batch.images, s.images), (batch.targets, s.targets)))
            if not same_bucket or batch.max_dec_len != s.max_dec_len:
                raise ValueError("step_graphed: padded shapes differ from the captured batch (another bucket); call capture() again")
            # any lengths of the same bucket: the graph reads them from the static plans (datasets.py:298-302 pads every batch to
            # its own maxima, so a training run needs one capture per padded-shape bucket, not per batch)
            for plan, lens in zip(self._plans, (batch.text_len, batch.audio_len, batch.image_len)):
                plan.update(lens)
            s.text.copy_(batch.text, non_blocking=True)
            s.audio.copy_(batch.audio, non_blocking=True)
            s.images.copy_(batch.images, non_blocking=True)
            s.targets.copy_(batch.targets, non_blocking=True)
        self._graph.replay()
        self.grads.all_reduce_sum(self.group)
        self._graph_update.replay()
        return self._static_loss

    def _forward_backward(self, batch: Batch) -> torch.Tensor:
        self.model.train()
        if batch is getattr(self, "_static", None):
            for plan in self._plans:                     # inside the captured region: lengths -> int32, order, masks
                plan.refresh()
        _, loss = self.model(batch.text, batch.text_len, batch.audio, batch.audio_len, batch.images, batch.image_len,
                             batch.targets, batch.target_len, batch.max_dec_len)
        # functional backward: gradients are produced fresh and packed into the flat buffer with one copy
        # kernel (no per-parameter accumulation nodes, which also keeps the step CUDA-graph capturable)
        # (weight-gradient products run on side streams next to the serial chains; joined when the block exits)
        with leaf_lanes():
            grads = torch.autograd.grad(loss, self.grads.params, allow_unused=True)
        self.grads.pack(grads)
        return loss.detach()

    def _update(self) -> None:
        if self.grads.flat.is_cuda:
            # clip + Adadelta as one pass over the flat buffers (csrc/optimizer.cu); the norm stays a device scalar
            from . import ops
            o = self.optimizer
            self.last_grad_norm = torch.linalg.vector_norm(self.grads.flat)
            ops.adadelta_clip_step(self.grads.flat_param, self.grads.flat, o.square_avg, o.acc_delta, self.last_grad_norm,
                                   self.max_grad_norm, o.lr, o.rho, o.eps, o.wd)
            return
        self.last_grad_norm = self.grads.clip_(self.max_grad_norm)        # host tensors (gloo tests): the same update in torch ops
        self.optimizer.step()

    def step(self, batch: Batch) -> torch.Tensor:
        """forward + backward + all-reduce + clip + Adadelta on this rank's shard; returns the local loss."""
        loss = self._forward_backward(batch)
        self.grads.all_reduce_sum(self.group)
        self._update()
        return loss
