"""Host side of SURVEY.md 8f rank 4: the reference's on-disk feature formats -> padded batches in pinned memory.

What the reference's datasets.py reads, restated without its directory walking, index pickles and nltk matching:

* text   : a ``.pt`` file holding a dict  sentence -> 300-d tensor, in transcript order; the loader appends the all -1
           end-of-summary row and counts it in the length (``TextDataset.__getitem__``, datasets.py:63-70);
* audio  : a pickled numpy array ``(128, T)`` of MFCC features, transposed to ``(T, 128)`` (``AudioDataset.__getitem__``,
           datasets.py:183-188);
* targets: per video the float rows ``[idx_0], [idx_1], ..., [n_source]`` -- the selected sentence indices followed by the index of
           the EOS row (``TargetDataset.__getitem__``, datasets.py:269, :283-285); the sentence matching in front of it
           (nltk tokenisers over the ground-truth text) is not part of the hot path: indices come in as lists;
* images : the reference decodes JPEG key-frames and runs a frozen ResNet-101 per step (datasets.py:133-140, encoding.py:111-154);
           the path here starts at its 1000-d output, so key-frame features are a ``.pt`` tensor ``(Li, 1000)`` per video (our format).

Batches are padded exactly as the reference's ``collator`` / ``target_collator`` do (``pad_sequence(batch_first=True,
padding_value=0)`` + Python length lists, datasets.py:298-309) and assembled directly in page-locked buffers, so that the
host -> device copy of a step is asynchronous (``Batch.to(device, non_blocking=True)``, what ``bench.py``'s e2e arm times).
"""
from __future__ import annotations

import pickle
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .synth import Batch

EOS_VALUE = -1.0          # datasets.py:69


def load_text_features(path: str) -> Tuple[torch.Tensor, int]:
    """(n_sentences + 1, 300) with the EOS row last, and that length (datasets.py:63-70)."""
    emb = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(emb, dict) or not emb:
        raise ValueError(f"{path}: expected a non-empty dict sentence -> embedding")
    rows = list(emb.values())
    width = rows[0].numel()
    out = torch.empty(len(rows) + 1, width, dtype=torch.float32)
    for i, r in enumerate(rows):
        out[i] = r.reshape(width)
    out[len(rows)] = EOS_VALUE
    return out, len(rows) + 1


def load_audio_features(path: str) -> Tuple[torch.Tensor, int]:
    """(T, 128) from the pickled (128, T) array, and T (datasets.py:183-188).  The file is a pickle: trusted input only."""
    with open(path, "rb") as f:
        arr = pickle.load(f)
    if not isinstance(arr, np.ndarray) or arr.ndim != 2:
        raise ValueError(f"{path}: expected a pickled 2-D numpy array (features, frames)")
    t = torch.from_numpy(np.transpose(arr))
    return t, int(t.size(0))


def load_image_features(path: str) -> Tuple[torch.Tensor, int]:
    """(Li, 1000) key-frame features (the frozen CNN's output, see the module docstring), and Li."""
    t = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(t, torch.Tensor) or t.dim() != 2:
        raise ValueError(f"{path}: expected a (key-frames, features) tensor")
    return t.float(), int(t.size(0))


def target_rows(selected: Sequence[int], n_source_sentences: int) -> torch.Tensor:
    """(n + 1, 1) float rows: the selected sentence indices, then the EOS row index (datasets.py:269, :283-285)."""
    return torch.tensor([[float(i)] for i in selected] + [[float(n_source_sentences)]], dtype=torch.float32)


def _buffer(shape, dtype, pin: bool) -> torch.Tensor:
    return torch.zeros(shape, dtype=dtype, pin_memory=pin)


def collate(items: Sequence[torch.Tensor], pin: Optional[bool] = None) -> Tuple[torch.Tensor, List[int]]:
    """datasets.py:298-302 ``collator``: zero-padded (B, L_max, ...) + lengths; assembled in place in a (pinned) buffer."""
    if not items:
        raise ValueError("collate: empty batch")
    pin = torch.cuda.is_available() if pin is None else pin
    lengths = [int(t.size(0)) for t in items]
    out = _buffer((len(items), max(lengths)) + tuple(items[0].shape[1:]), items[0].dtype, pin)
    for b, t in enumerate(items):
        out[b, :lengths[b]] = t
    return out, lengths


def collate_targets(items: Sequence[torch.Tensor], pin: Optional[bool] = None) -> Tuple[torch.Tensor, List[int]]:
    """datasets.py:304-309 ``target_collator`` (the two path tuples it also returns are bookkeeping): (B, T_max, 1) + lengths."""
    return collate(items, pin)


def load_batch(text_paths: Iterable[str], audio_paths: Iterable[str], image_paths: Iterable[str],
               selected: Sequence[Sequence[int]], pin: Optional[bool] = None) -> Batch:
    """One training batch from feature files: what ``train.py:125-135`` zips out of its four DataLoaders, as a ``Batch``
    (images shaped (B, Li, E, 1, 1) as ``MMBiDAF.forward`` takes them; ``max_dec_len`` = the longest target, evaluate.py:113)."""
    texts = [load_text_features(p) for p in text_paths]
    audios = [load_audio_features(p)[0] for p in audio_paths]
    images = [load_image_features(p)[0] for p in image_paths]
    if not (len(texts) == len(audios) == len(images) == len(selected)):
        raise ValueError("load_batch: the four modalities must list the same videos")
    text, text_len = collate([t for t, _ in texts], pin)
    audio, audio_len = collate(audios, pin)
    image, image_len = collate(images, pin)
    targets, target_len = collate_targets([target_rows(sel, n - 1) for sel, (_, n) in zip(selected, texts)], pin)
    return Batch(text, text_len, audio, audio_len, image.unsqueeze(-1).unsqueeze(-1), image_len, targets, target_len,
                 max(target_len))


# ---------------------------------------------------------------------------------------------------------------------
# Packed layout.  The reference's transcript files are dicts of ~400 separately pickled 300-d tensors: un-pickling one costs
# ~20 ms (tools/data_loader_bench.py: 28 videos/s per host thread at BASELINE config 3 sizes, against ~6 000 videos/s for the
# training step on one B200).  Converted once, every modality is ONE row-major float32 matrix of all videos' rows plus an offset
# table, memory-mapped: a video is a contiguous slice, a batch is B slice copies straight into the pinned buffer.
# ---------------------------------------------------------------------------------------------------------------------
def pack_features(items: Iterable[torch.Tensor], prefix: str) -> None:
    """Write ``prefix.rows.npy`` ((sum of rows, width) float32) and ``prefix.offsets.npy`` ((n + 1,) int64)."""
    mats = [np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32) for t in items]
    if not mats or any(m.ndim != 2 or m.shape[1] != mats[0].shape[1] for m in mats):
        raise ValueError("pack_features: need a non-empty list of (rows, width) matrices of one width")
    offsets = np.zeros(len(mats) + 1, dtype=np.int64)
    np.cumsum([m.shape[0] for m in mats], out=offsets[1:])
    np.save(prefix + ".rows.npy", np.concatenate(mats, axis=0))
    np.save(prefix + ".offsets.npy", offsets)


class PackedFeatures:
    """Memory-mapped view of what :func:`pack_features` wrote: ``len()`` videos, ``[i]`` -> (rows_i, width) tensor (a copy;
    ``collate`` copies slices of the map straight into the batch buffer)."""

    def __init__(self, prefix: str):
        self.rows = np.load(prefix + ".rows.npy", mmap_mode="r")
        self.offsets = np.load(prefix + ".offsets.npy")
        if self.rows.ndim != 2 or self.offsets.ndim != 1 or int(self.offsets[-1]) != self.rows.shape[0]:
            raise ValueError(f"{prefix}: rows / offsets do not match")

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def length(self, i: int) -> int:
        return int(self.offsets[i + 1] - self.offsets[i])

    def __getitem__(self, i: int) -> torch.Tensor:
        return torch.from_numpy(np.array(self.rows[int(self.offsets[i]):int(self.offsets[i + 1])]))     # a private copy

    def collate(self, indices: Sequence[int], pin: Optional[bool] = None) -> Tuple[torch.Tensor, List[int]]:
        """The reference collator (datasets.py:298-302) over videos ``indices``, slices copied straight into the batch buffer."""
        if not len(indices):
            raise ValueError("collate: empty batch")
        pin = torch.cuda.is_available() if pin is None else pin
        lengths = [self.length(i) for i in indices]
        out = _buffer((len(indices), max(lengths), self.rows.shape[1]), torch.float32, pin)
        dst = out.numpy()
        for b, i in enumerate(indices):
            dst[b, :lengths[b]] = self.rows[int(self.offsets[i]):int(self.offsets[i + 1])]
        return out, lengths


def pack_videos(text_paths: Sequence[str], audio_paths: Sequence[str], image_paths: Sequence[str], prefix: str) -> None:
    """Convert feature files in the reference's formats (see the module docstring) to the packed layout, once."""
    pack_features((load_text_features(p)[0] for p in text_paths), prefix + ".text")
    pack_features((load_audio_features(p)[0] for p in audio_paths), prefix + ".audio")
    pack_features((load_image_features(p)[0] for p in image_paths), prefix + ".image")


class PackedDataset:
    """The three packed modalities of a set of videos; ``batch(indices, selected)`` == :func:`load_batch` on the same videos."""

    def __init__(self, prefix: str):
        self.text, self.audio, self.image = (PackedFeatures(prefix + s) for s in (".text", ".audio", ".image"))
        if not (len(self.text) == len(self.audio) == len(self.image)):
            raise ValueError(f"{prefix}: the modalities list different numbers of videos")

    def __len__(self) -> int:
        return len(self.text)

    def batch(self, indices: Sequence[int], selected: Sequence[Sequence[int]], pin: Optional[bool] = None) -> Batch:
        if len(indices) != len(selected):
            raise ValueError("batch: one list of selected sentence indices per video")
        text, text_len = self.text.collate(indices, pin)
        audio, audio_len = self.audio.collate(indices, pin)
        image, image_len = self.image.collate(indices, pin)
        targets, target_len = collate_targets([target_rows(sel, n - 1) for sel, n in zip(selected, text_len)], pin)
        return Batch(text, text_len, audio, audio_len, image.unsqueeze(-1).unsqueeze(-1), image_len, targets, target_len,
                     max(target_len))
