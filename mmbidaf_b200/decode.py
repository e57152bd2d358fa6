"""Extractive decoding over the model's output distributions (reference evaluate.py:167-202)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def greedy_search(out_distributions: torch.Tensor, original_text_length: int) -> List[int]:
    """Sentence indices picked for one video: arg-max per timestep until it hits the EOS row
    ``original_text_length - 1`` (evaluate.py:185-202).  The reference looks every pick up in the transcript on disk
    (``get_source_sentence``, :236-259): an index past the transcript yields no sentence and is skipped, not a stop
    (:195, :255-256) -- reproduced here from the length alone (the transcript has ``original_text_length - 1`` sentences)."""
    picks = out_distributions.argmax(dim=1).tolist()
    chosen = []
    for k in picks:
        if k == original_text_length - 1:
            break
        if k > original_text_length - 1:
            continue
        chosen.append(int(k))
    return chosen


def get_generated_indices(batch_out_distributions: torch.Tensor, original_text_lengths: Sequence[int]) -> List[List[int]]:
    """Greedy indices for a batch (B, T, M) (evaluate.py:167-183, method='greedy')."""
    host = batch_out_distributions.detach().cpu()
    return [greedy_search(host[b], int(original_text_lengths[b])) for b in range(host.shape[0])]
