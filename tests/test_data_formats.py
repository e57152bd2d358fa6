"""CPU: the on-disk feature formats and collators against fixtures made by the reference's own
datasets.py (tests/golden/make_golden_data.py): bit-exact."""
import os

import pytest
import torch

from conftest import ROOT
from mmbidaf_b200 import data

DATA = os.path.join(ROOT, "tests", "golden", "data")


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "data_formats.pt"), weights_only=True)


def test_text_and_audio_items_match_the_reference_datasets(golden):
    for i, (want, n) in enumerate(zip(golden["text_items"], golden["text_item_len"])):
        got, got_n = data.load_text_features(os.path.join(DATA, f"text_{i}.pt"))
        assert got_n == n and got.dtype == want.dtype and torch.equal(got, want)
        assert torch.equal(got[-1], torch.full((300,), -1.0))                 # the EOS row (datasets.py:69)
    for i, (want, n) in enumerate(zip(golden["audio_items"], golden["audio_item_len"])):
        got, got_n = data.load_audio_features(os.path.join(DATA, f"audio_{i}.pkl"))
        assert got_n == n and got.dtype == want.dtype and torch.equal(got, want)


def test_collators_match_the_reference(golden):
    text, text_len = data.collate(golden["text_items"], pin=False)
    audio, audio_len = data.collate(golden["audio_items"], pin=False)
    assert text_len == golden["text_len"] and torch.equal(text, golden["text_padded"])
    assert audio_len == golden["audio_len"] and torch.equal(audio, golden["audio_padded"])
    n_source = [n - 1 for n in golden["text_item_len"]]
    rows = [data.target_rows(sel, n) for sel, n in zip(golden["picks"], n_source)]
    targets, target_len = data.collate_targets(rows, pin=False)
    assert target_len == golden["target_len"] and targets.dtype == golden["targets_padded"].dtype
    assert torch.equal(targets, golden["targets_padded"])


def test_load_batch_assembles_a_model_batch(golden, tmp_path):
    gen = torch.Generator().manual_seed(1)
    image_paths = []
    for i, li in enumerate((4, 2, 3)):
        p = tmp_path / f"img_{i}.pt"
        torch.save(torch.randn(li, 1000, generator=gen), p)
        image_paths.append(str(p))
    b = data.load_batch([os.path.join(DATA, f"text_{i}.pt") for i in range(3)],
                        [os.path.join(DATA, f"audio_{i}.pkl") for i in range(3)], image_paths, golden["picks"], pin=False)
    assert torch.equal(b.text, golden["text_padded"]) and b.text_len == golden["text_len"]
    assert torch.equal(b.audio, golden["audio_padded"]) and b.audio_len == golden["audio_len"]
    assert b.images.shape == (3, 4, 1000, 1, 1) and b.image_len == [4, 2, 3] and float(b.images[1, 2:].abs().sum()) == 0.0
    assert torch.equal(b.targets, golden["targets_padded"]) and b.target_len == golden["target_len"] and b.max_dec_len == 5
    # the last valid target of every video is its EOS row (datasets.py:283), which is the last valid text row
    for v in range(3):
        assert int(b.targets[v, b.target_len[v] - 1, 0]) == b.text_len[v] - 1


def test_bad_inputs_raise(tmp_path):
    p = tmp_path / "not_a_dict.pt"
    torch.save(torch.zeros(3), p)
    with pytest.raises(ValueError):
        data.load_text_features(str(p))
    with pytest.raises(ValueError):
        data.collate([])
    with pytest.raises(ValueError):
        data.load_batch([], [], [], [[0]])


def test_packed_layout_round_trips_bit_exactly(golden, tmp_path):
    gen = torch.Generator().manual_seed(2)
    image_paths = []
    for i, li in enumerate((4, 2, 3)):
        p = tmp_path / f"img_{i}.pt"
        torch.save(torch.randn(li, 1000, generator=gen), p)
        image_paths.append(str(p))
    texts = [os.path.join(DATA, f"text_{i}.pt") for i in range(3)]
    audios = [os.path.join(DATA, f"audio_{i}.pkl") for i in range(3)]
    prefix = str(tmp_path / "packed")
    data.pack_videos(texts, audios, image_paths, prefix)
    ds = data.PackedDataset(prefix)
    assert len(ds) == 3 and torch.equal(ds.text[1], golden["text_items"][1]) and torch.equal(ds.audio[2], golden["audio_items"][2])
    want = data.load_batch(texts, audios, image_paths, golden["picks"], pin=False)
    got = ds.batch([0, 1, 2], golden["picks"], pin=False)
    for name in ("text", "audio", "images", "targets"):
        assert torch.equal(getattr(got, name), getattr(want, name)), name
    assert (got.text_len, got.audio_len, got.image_len, got.target_len, got.max_dec_len) == \
           (want.text_len, want.audio_len, want.image_len, want.target_len, want.max_dec_len)
    sub = ds.batch([2, 0], [golden["picks"][2], golden["picks"][0]], pin=False)          # any order / subset
    assert sub.text_len == [8, 6] and torch.equal(sub.text[1, :6], golden["text_items"][0])
    with pytest.raises(ValueError):
        ds.batch([0], [], pin=False)
