"""Host-side measurement for SURVEY.md 8f rank 4: feature files in the reference's formats -> one padded (pinned) batch.
Writes `--videos` synthetic videos at BASELINE config 3 sizes to a temporary directory, then times mmbidaf_b200.data.load_batch
and -- when the reference tree is present (build container only) -- the reference's own TextDataset / AudioDataset __getitem__
+ collators on the same files (nltk stubbed as in tests/golden/make_golden_data.py).
    python tools/data_loader_bench.py [--videos 32] [--reps 5]
"""
import argparse
import os
import pickle
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmbidaf_b200 import data  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=32)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
gen = torch.Generator().manual_seed(224)
tmp = tempfile.mkdtemp(prefix="mmb_data_")
texts, audios, images, picks = [], [], [], []
for v in range(a.videos):
    n_sent = int(torch.randint(204, 409, (1,), generator=gen))
    frames = int(torch.randint(512, 1025, (1,), generator=gen))
    li = int(torch.randint(64, 129, (1,), generator=gen))
    torch.save({f"sentence {k} of video {v}": torch.randn(300, generator=gen) for k in range(n_sent)}, f"{tmp}/text_{v}.pt")
    with open(f"{tmp}/audio_{v}.pkl", "wb") as f:
        pickle.dump(torch.randn(128, frames, generator=gen).numpy().astype(np.float32), f, protocol=4)
    torch.save(torch.randn(li, 1000, generator=gen), f"{tmp}/img_{v}.pt")
    texts.append(f"{tmp}/text_{v}.pt"); audios.append(f"{tmp}/audio_{v}.pkl"); images.append(f"{tmp}/img_{v}.pt")
    picks.append(torch.randint(0, n_sent, (11,), generator=gen).tolist())


def timed(fn):
    fn()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        out = fn()
    return (time.perf_counter() - t0) / a.reps, out


pin = torch.cuda.is_available()
t_ours, batch = timed(lambda: data.load_batch(texts, audios, images, picks, pin=pin))
print(f"mmbidaf_b200.data.load_batch: {a.videos} videos in {t_ours * 1e3:.1f} ms = {a.videos / t_ours:.0f} videos/s on one host thread "
      f"({batch.h2d_bytes() / 1e6:.1f} MB, pinned={pin})")
t0 = time.perf_counter()
data.pack_videos(texts, audios, images, f"{tmp}/packed")
t_pack = time.perf_counter() - t0
ds = data.PackedDataset(f"{tmp}/packed")
t_packed, batch2 = timed(lambda: ds.batch(list(range(a.videos)), picks, pin=pin))
assert all(torch.equal(getattr(batch, n), getattr(batch2, n)) for n in ("text", "audio", "images", "targets"))
print(f"packed layout (one-off conversion {t_pack * 1e3:.0f} ms): {a.videos} videos in {t_packed * 1e3:.2f} ms = "
      f"{a.videos / t_packed:.0f} videos/s on one host thread, {batch.h2d_bytes() / t_packed / 1e9:.1f} GB/s, identical tensors")
ref_root = os.environ.get("MMBIDAF_REFERENCE", "/root/reference")
if os.path.isdir(ref_root):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
    from make_golden_data import import_reference_datasets
    ref = import_reference_datasets()
    tds, ads = object.__new__(ref.TextDataset), object.__new__(ref.AudioDataset)
    tds.text_embedding_paths, ads.audios_paths = texts, audios

    def reference():
        t, tl = ref.collator([tds[i] for i in range(a.videos)])
        au, al = ref.collator([ads[i] for i in range(a.videos)])
        return t, au

    t_ref, (t, au) = timed(reference)
    assert torch.equal(t, batch.text) and torch.equal(au, batch.audio)
    t_two, _ = timed(lambda: (data.collate([data.load_text_features(p)[0] for p in texts], pin=pin),
                              data.collate([data.load_audio_features(p)[0] for p in audios], pin=pin)))
    print(f"text + audio only: reference datasets.py {t_ref * 1e3:.1f} ms, mmbidaf_b200.data {t_two * 1e3:.1f} ms (identical tensors)")
