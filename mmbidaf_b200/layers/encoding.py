"""Encoders of the MMBiDAF hot path on B200.

Same public surface as the reference's layers/encoding.py (``Embedding``, ``HighwayEncoder``,
``RNNEncoder``, ``ImageEmbedding``): identical constructor and ``forward`` signatures and identical
parameter names / shapes, so a reference checkpoint loads with ``load_state_dict``.  The recurrent
part runs in the persistent LSTM kernels of csrc/bilstm.cu; there is no cuDNN / CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops

__all__ = ["Embedding", "HighwayEncoder", "RNNEncoder", "ImageEmbedding"]


class HighwayEncoder(nn.Module):
    """``num_layers`` highway layers x <- g*relu(T x) + (1-g)*x  (reference encoding.py:45-59).
    Each layer is one GEMM (gate and transform stacked) plus one fused point-wise kernel, forward and backward
    (csrc/highway.cu); the reference runs two GEMMs and seven element-wise kernels per layer."""

    def __init__(self, num_layers, hidden_size):
        super().__init__()
        self.transforms = nn.ModuleList(nn.Linear(hidden_size, hidden_size) for _ in range(num_layers))
        self.gates = nn.ModuleList(nn.Linear(hidden_size, hidden_size) for _ in range(num_layers))

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("mmbidaf_b200.layers.HighwayEncoder runs on a B200 only (no CPU fallback)")
        for gate, transform in zip(self.gates, self.transforms):
            x = Fn.highway_layer(x, gate, transform)
        return x


class Embedding(nn.Module):
    """dropout -> Linear(E -> H, no bias) -> 2-layer highway  (reference encoding.py:19-30)."""

    def __init__(self, embedding_size, hidden_size, drop_prob):
        super().__init__()
        self.drop_prob = drop_prob
        self.proj = nn.Linear(embedding_size, hidden_size, bias=False)
        self.hwy = HighwayEncoder(2, hidden_size)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("mmbidaf_b200.layers.Embedding runs on a B200 only (no CPU fallback)")
        x = Fn.dropout(x, self.drop_prob, self.training)         # encoding.py:26, one own launch (counter-based keep bits)
        return self.hwy(Fn.tall_linear(x, self.proj.weight))     # Linear(E -> H, no bias); batched weight gradient


class LengthPlan:
    """Device-side view of ONE host list of lengths (the reference passes lengths as Python lists, datasets.py:300-302).

    ``packed`` (2, B) int64 = [lengths, sort index]; the sort index is computed exactly as the reference does (encoding.py:85,91:
    a CPU float tensor sorted with torch.sort descending) because the rows of the returned hidden state keep that order (quirk
    Q3) and that sort's tie order is not reproducible elsewhere (it is not stable; csrc/length_plan.cu).  Everything else is made
    on the device from the lengths by one small kernel per mask shape (``ops.length_plan``): int32 lengths, the longest-first
    scheduling order of the recurrences, position masks and the decoder mask (models.py:86-92, :119-123).

    A *static* plan keeps every device tensor at a fixed address: ``update`` rewrites the lengths (pinned host ring -> H2D) and
    ``refresh`` re-launches the kernels, so a captured CUDA graph that contains the ``refresh`` launches serves any batch."""

    RING = 4

    def __init__(self, lengths, device, static: bool = False):
        n = len(lengths)
        self.device, self.static = device, static
        self.packed = torch.empty(2, n, dtype=torch.int64, device=device)
        self.len_i32 = torch.empty(n, dtype=torch.int32, device=device)
        self.order_i32 = torch.empty(n, dtype=torch.int32, device=device)
        self.masks = {}
        self._host = [torch.empty(2, n, dtype=torch.int64).pin_memory() for _ in range(self.RING)] if static else None
        self._events = [None] * self.RING
        self._turn = 0
        self.update(lengths)
        self.refresh()

    @property
    def sort_idx(self):
        return self.packed[1]

    def update(self, lengths) -> None:
        """Host side of a new list of lengths: the reference's own sort, then one small H2D copy (stream ordered)."""
        vals = [int(v) for v in lengths]
        if len(vals) != self.packed.shape[1]:
            raise ValueError(f"LengthPlan.update: {len(vals)} lengths for a plan of {self.packed.shape[1]}")
        sort_idx = torch.Tensor(vals).sort(0, descending=True)[1]                 # encoding.py:85, :91
        if self.static:
            slot = self._turn % self.RING
            self._turn += 1
            if self._events[slot] is not None:
                self._events[slot].synchronize()                                   # the copy that last read this pinned slot
            host = self._host[slot]
            host[0] = torch.tensor(vals, dtype=torch.int64)
            host[1] = sort_idx
            self.packed.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._events[slot] = ev
        else:
            self.packed.copy_(torch.stack([torch.tensor(vals, dtype=torch.int64), sort_idx]), non_blocking=True)

    def refresh(self) -> None:
        """Device side: everything derived from the lengths, into the same tensors (capturable)."""
        from .. import ops
        self.len_i32.copy_(self.packed[0])
        ops.length_plan(self.len_i32, 0, 0, out=(None, None, self.order_i32))
        for (L, M), (mask, dec) in self.masks.items():
            ops.length_plan(self.len_i32, L, M, out=(mask, dec, None))

    def mask(self, L: int, M: int = 0):
        """(mask (B, L) bool, decoder mask (B, M) bool or None); made once per shape, re-made by ``refresh``."""
        from .. import ops
        hit = self.masks.get((L, M))
        if hit is None:
            mask, dec, _ = ops.length_plan(self.len_i32, L, M)
            hit = self.masks[(L, M)] = (mask, dec)
        return hit


class _LengthCache:
    """Host list of lengths -> LengthPlan.  Plans are cached per distinct list of values; ``pin`` registers a static plan for one
    particular list OBJECT (looked up by identity first), which is how a graph-captured step keeps its length tensors."""

    def __init__(self):
        self._store = {}
        self._pins = {}

    def get(self, lengths, device) -> LengthPlan:
        pin = self._pins.get(id(lengths))
        if pin is not None and pin[0] is lengths:
            return pin[1]
        key = (tuple(int(v) for v in lengths), str(device))
        hit = self._store.get(key)
        if hit is None:
            if len(self._store) > 64:
                self._store.clear()
            hit = self._store[key] = LengthPlan(key[0], device)
        return hit

    def pin(self, lengths, device) -> LengthPlan:
        pin = self._pins.get(id(lengths))
        if pin is None or pin[0] is not lengths:
            pin = self._pins[id(lengths)] = (lengths, LengthPlan(lengths, device, static=True))
        return pin[1]

    def unpin(self, lengths) -> None:
        pin = self._pins.get(id(lengths))
        if pin is not None and pin[0] is lengths:
            del self._pins[id(lengths)]


_lengths = _LengthCache()


def length_plan(lengths, device) -> LengthPlan:
    return _lengths.get(lengths, device)


def pin_lengths(lengths, device) -> LengthPlan:
    """A static plan for this list object (see LengthPlan); the caller keeps the returned plan alive."""
    return _lengths.pin(lengths, device)


def unpin_lengths(lengths) -> None:
    _lengths.unpin(lengths)


def device_lengths(lengths, device):
    """int32 device tensor of a host list of lengths (cached per distinct list)."""
    return _lengths.get(lengths, device).len_i32


class RNNEncoder(nn.Module):
    """Length-aware bidirectional LSTM encoder (reference encoding.py:62-108).

    ``self.rnn`` is an ``nn.LSTM`` used purely as the parameter container (names
    ``rnn.weight_ih_l{k}[_reverse]`` ... and default init identical to the reference); its own forward is
    never called.  No sort / pack / unpack gathers: the kernel walks each sample up to its own length and
    writes exact zeros past it.  Returns ``(x, x_hidden)`` with ``x_hidden`` (B, 2*layers, H) left in
    descending-length row order, as the reference leaves it (quirk Q3, encoding.py:99-106)."""

    def __init__(self, input_size, hidden_size, num_layers, drop_prob=0.):
        super().__init__()
        self.drop_prob = drop_prob
        self.num_layers = num_layers
        self.rnn = nn.LSTM(input_size, hidden_size, num_layers, batch_first=True, bidirectional=True,
                           dropout=drop_prob if num_layers > 1 else 0.)

    def _layer_weights(self, k):
        return [getattr(self.rnn, f"{kind}_l{k}{suffix}") for suffix in ("", "_reverse")
                for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    def forward(self, x, lengths):
        if not x.is_cuda:
            raise RuntimeError("mmbidaf_b200.layers.RNNEncoder runs on a B200 only (no CPU fallback)")
        plan = _lengths.get(lengths, x.device)
        finals = []
        # Dropout of every layer's output -- nn.LSTM's inter-layer dropout (encoding.py:77-81) and F.dropout on the encoder's
        # output (encoding.py:104) -- happens inside the recurrence kernels: one launch draws the layer keys on the device.
        p_inter = self.rnn.dropout if self.training else 0.0
        p_out = self.drop_prob if self.training else 0.0
        if p_out >= 1.0 or p_inter >= 1.0:
            raise ValueError("RNNEncoder: drop_prob must be < 1")
        keys = ops.rng_next_keys(x.device, self.num_layers) if (p_inter > 0.0 or p_out > 0.0) else None
        for k in range(self.num_layers):
            p = p_inter if k + 1 < self.num_layers else p_out
            x, h_n = Fn.lstm_layer(x, plan.len_i32, plan.order_i32, self._layer_weights(k),
                                   None if keys is None else keys[k:k + 1], p)
            finals.append(h_n)
        x_hidden = torch.cat(finals, dim=1).index_select(0, plan.sort_idx)
        return x, x_hidden


class ImageEmbedding(nn.Module):
    """Key-frame encoder (reference encoding.py:111-154): a frozen ImageNet ResNet-101 giving 1000-d
    logits per frame.  The CNN is outside the hot path; it is built lazily on first use with real images
    so that constructing the model needs no weight download.  Pre-extracted features shaped
    (N, E, 1, 1) pass straight through (north_star feeds "image 1000-d")."""

    def __init__(self):
        super().__init__()
        self.resnet = None                 # registered lazily under the reference's attribute name

    def _build(self, pretrained, device=None):
        import torchvision
        weights = torchvision.models.ResNet101_Weights.IMAGENET1K_V1 if pretrained else None
        net = torchvision.models.resnet101(weights=weights)
        for p in net.parameters():         # fine_tune(False), reference encoding.py:141-154
            p.requires_grad = False
        self.resnet = net.to(device) if device is not None else net
        return self.resnet

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # a reference checkpoint carries image_keyframes_emb.resnet.*: materialise the CNN to receive it
        if self.resnet is None and any(k.startswith(prefix + "resnet.") for k in state_dict):
            self._build(pretrained=False)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, images):
        if images.dim() == 4 and images.shape[2] == 1 and images.shape[3] == 1:
            return images.flatten(1)
        net = self.resnet if self.resnet is not None else self._build(pretrained=True, device=images.device)
        return net(images)
