"""Golden fixtures for the on-disk feature formats, made FROM THE REFERENCE's own datasets.py.

Runs only in the build container (reference mounted read-only at /root/reference).  Writes small feature files in the
reference's formats under tests/golden/data/ -- a ``.pt`` dict sentence -> 300-d tensor per transcript (read by
``TextDataset.__getitem__``, datasets.py:63-70), a pickled ``(128, T)`` numpy array per audio track (``AudioDataset.__getitem__``,
datasets.py:183-188) -- then runs the reference's ``__getitem__`` methods and collators (datasets.py:298-309) on them and stores
what they return in tests/golden/data_formats.pt.  ``datasets.py`` imports nltk (absent here, used only by the target-sentence
matching, which is out of scope): it is stubbed in ``sys.modules`` for the import; the dataset constructors walk a course
directory tree and pickled index files that are not shipped, so instances are made with ``object.__new__`` and given the path
lists directly.  Nothing at test time reads /root/reference.

    python tests/golden/make_golden_data.py
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")
REF = os.environ.get("MMBIDAF_REFERENCE", "/root/reference")


def import_reference_datasets():
    for name in ("nltk", "nltk.corpus", "nltk.tokenize", "nltk.stem"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["nltk.corpus"].stopwords = sys.modules["nltk.corpus"].words = None
    for attr in ("sent_tokenize", "word_tokenize", "TweetTokenizer"):
        setattr(sys.modules["nltk.tokenize"], attr, None)
    sys.modules["nltk.stem"].WordNetLemmatizer = None
    sys.path.insert(0, REF)
    sys.modules.pop("datasets", None)
    import datasets as ref_datasets
    return ref_datasets


def write_inputs():
    os.makedirs(DATA, exist_ok=True)
    gen = torch.Generator().manual_seed(224)
    texts, audios = [], []
    for i, n_sent in enumerate((5, 3, 7)):
        emb = {f"sentence {i}-{k} of the transcript": torch.randn(300, generator=gen) for k in range(n_sent)}
        path = os.path.join(DATA, f"text_{i}.pt")
        torch.save(emb, path)
        texts.append(path)
    for i, frames in enumerate((9, 4, 6)):
        mfcc = torch.randn(128, frames, generator=gen).numpy().astype(np.float32)       # (128, T), as the reference stores it
        path = os.path.join(DATA, f"audio_{i}.pkl")
        with open(path, "wb") as f:
            pickle.dump(mfcc, f, protocol=4)
        audios.append(path)
    return texts, audios


def main():
    ref = import_reference_datasets()
    texts, audios = write_inputs()
    text_ds = object.__new__(ref.TextDataset)
    text_ds.text_embedding_paths = texts
    audio_ds = object.__new__(ref.AudioDataset)
    audio_ds.audios_paths = audios
    text_items = [text_ds[i] for i in range(len(texts))]
    audio_items = [audio_ds[i] for i in range(len(audios))]
    text_padded, text_len = ref.collator(text_items)
    audio_padded, audio_len = ref.collator(audio_items)
    # targets: TargetDataset.__getitem__ (datasets.py:232-285) ends with stack([Tensor([idx]) ...] + [Tensor([n_source])]);
    # the sentence matching in front of it needs nltk and the transcripts' text and is out of scope
    picks = ([0, 2, 3], [1], [4, 0, 2, 5])
    n_source = [len(torch.load(p)) for p in texts]
    target_items = []
    for idx, (sel, n) in enumerate(zip(picks, n_source)):
        rows = [torch.Tensor([s]) for s in sel] + [torch.Tensor([n])]                     # datasets.py:269, :283
        target_items.append((torch.stack(rows), texts[idx], "ground-truth", len(rows)))   # datasets.py:285
    tgt_padded, src_paths, tgt_paths, tgt_len = ref.target_collator(target_items)
    out = {
        "text_items": [t for t, _ in text_items], "text_item_len": [n for _, n in text_items],
        "audio_items": [a for a, _ in audio_items], "audio_item_len": [n for _, n in audio_items],
        "text_padded": text_padded, "text_len": text_len, "audio_padded": audio_padded, "audio_len": audio_len,
        "picks": [list(p) for p in picks], "targets_padded": tgt_padded, "target_len": tgt_len,
        "torch": str(torch.__version__),
    }
    torch.save(out, os.path.join(HERE, "data_formats.pt"))
    print({k: tuple(v.shape) for k, v in out.items() if hasattr(v, "shape")}, out["text_len"], out["audio_len"], out["target_len"])


if __name__ == "__main__":
    main()
