"""Synthetic Coursera-shaped batches.

Shapes follow what the reference's collators emit (datasets.py:298-309): zero-padded
float feature tensors plus Python lists of lengths; the last valid text row is the
all -1 EOS row (datasets.py:69) and lengths include it (datasets.py:70); targets are a
float (B, T, 1) tensor whose last valid entry is the EOS row index (datasets.py:283),
zero padded (datasets.py:308).  Image key-frames are fed as pre-extracted 1000-d ResNet
logits shaped (B, Li, E, 1, 1): the frozen CNN is outside the hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import torch


@dataclass
class Batch:
    text: torch.Tensor            # (B, Lt, E_text)
    text_len: List[int]
    audio: torch.Tensor           # (B, La, E_audio)
    audio_len: List[int]
    images: torch.Tensor          # (B, Li, E_image, 1, 1)
    image_len: List[int]
    targets: torch.Tensor         # (B, T, 1) float
    target_len: List[int]
    max_dec_len: int

    def to(self, device, non_blocking: bool = False) -> "Batch":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return Batch(mv(self.text), self.text_len, mv(self.audio), self.audio_len, mv(self.images),
                     self.image_len, mv(self.targets), self.target_len, self.max_dec_len)

    def pin(self) -> "Batch":
        pn = lambda t: t.pin_memory()
        return Batch(pn(self.text), self.text_len, pn(self.audio), self.audio_len, pn(self.images),
                     self.image_len, pn(self.targets), self.target_len, self.max_dec_len)

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.text, self.audio, self.images, self.targets))


def ragged_lengths(batch: int, max_len: int, gen: torch.Generator, min_len: int = 1) -> List[int]:
    """~U{ceil(L/2)..L}; sample 0 forced to L so the padded width equals L."""
    lo = max(min_len, math.ceil(max_len / 2))
    lens = torch.randint(lo, max_len + 1, (batch,), generator=gen).tolist()
    lens[0] = max_len
    return lens


def padded_features(lengths: List[int], max_len: int, width: int, gen: torch.Generator) -> torch.Tensor:
    x = torch.randn(len(lengths), max_len, width, generator=gen)
    for b, n in enumerate(lengths):
        x[b, n:] = 0.0
    return x


def make_batch(batch: int, lt: int, la: int, li: int, t_dec: int, e_text: int = 300, e_audio: int = 128,
               e_image: int = 1000, seed: int = 224) -> Batch:
    gen = torch.Generator().manual_seed(seed)
    text_len = ragged_lengths(batch, lt, gen, min_len=2)
    audio_len = ragged_lengths(batch, la, gen)
    image_len = ragged_lengths(batch, li, gen)
    text = padded_features(text_len, lt, e_text, gen)
    for b, n in enumerate(text_len):
        text[b, n - 1] = -1.0                                     # EOS row
    audio = padded_features(audio_len, la, e_audio, gen)
    images = padded_features(image_len, li, e_image, gen).reshape(batch, li, e_image, 1, 1)
    target_len = ragged_lengths(batch, t_dec, gen)
    targets = torch.zeros(batch, t_dec, 1)
    for b, n in enumerate(target_len):
        if n > 1:
            targets[b, : n - 1, 0] = torch.randint(0, max(text_len[b] - 1, 1), (n - 1,), generator=gen).float()
        targets[b, n - 1, 0] = float(text_len[b] - 1)             # EOS index
    return Batch(text, text_len, audio, audio_len, images, image_len, targets, target_len, max(target_len))
