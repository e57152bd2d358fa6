"""The staged reference (oracle/_ref, made by oracle/stage_ref.py from /root/reference, byte for byte) is the reference:
on the CPU it reproduces the committed golden fixtures, and its own ``models.py`` -- unmodified -- builds over
``mmbidaf_b200.layers`` with an identical parameter set (the drop-in contract of north_star / SURVEY.md section 8b).
No GPU compute here; the GPU half is tests/test_reference_dropin_gpu.py."""
import hashlib
import json
import os

import pytest
import torch

from conftest import ROOT, load_golden, rel_err
from mmbidaf_b200.synth import Batch
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref is not staged (python oracle/stage_ref.py)")


def test_manifest_matches_staged_bytes():
    with open(os.path.join(ref_loader.REF, "MANIFEST.json")) as f:
        manifest = json.load(f)
    assert {"models.py", "layers/attention.py", "layers/encoding.py"} <= set(manifest)
    for rel, digest in manifest.items():
        with open(os.path.join(ref_loader.REF, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
    # nothing under oracle/_ref is tracked by git
    ignore = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ignore


def test_staged_reference_reproduces_golden_on_cpu():
    g = load_golden("model_small.pt")
    hidden, e_t, e_a, e_i, m = g["dims"]
    torch.set_num_threads(1)
    model = ref_loader.build_model("reference", hidden, e_t, e_a, e_i, torch.device("cpu"), 0.0, m, params=g["params"])
    batch = Batch(**g["batch"])
    model.train()
    out, loss = model(batch.text, batch.text_len, batch.audio, batch.audio_len, batch.images, batch.image_len, batch.targets,
                      batch.target_len, batch.max_dec_len)
    assert rel_err(out, g["train_out"]) < 1e-6 and rel_err(loss, g["train_loss"]) < 1e-6
    loss.backward()
    for name, p in model.named_parameters():
        if name in g["train_grads"]:
            assert rel_err(p.grad, g["train_grads"][name]) < 1e-4 or float(g["train_grads"][name].abs().max()) < 1e-6, name
    model.eval()
    with torch.no_grad():
        out_e, _ = model(batch.text, batch.text_len, batch.audio, batch.audio_len, batch.images, batch.image_len, batch.targets,
                         batch.target_len, batch.max_dec_len)
    assert torch.equal(out_e.argmax(dim=2), g["eval_argmax"])


def test_reference_models_py_builds_over_b200_layers():
    """models.py:4-5 `from layers.X import *` resolving to mmbidaf_b200.layers: same modules, names and shapes."""
    g = load_golden("model_small.pt")
    hidden, e_t, e_a, e_i, m = g["dims"]
    ours = ref_loader.build_model("b200", hidden, e_t, e_a, e_i, torch.device("cpu"), 0.0, m, params=g["params"])
    ref = ref_loader.build_model("reference", hidden, e_t, e_a, e_i, torch.device("cpu"), 0.0, m, params=g["params"])
    import mmbidaf_b200.layers as L
    assert type(ours.bidaf_att_audio) is L.BiDAFAttention and type(ours.mod_t_a) is L.RNNEncoder
    assert type(ours.multimodal_att_decoder) is L.MultimodalAttentionDecoder and type(ours.emb) is L.Embedding
    a, b = dict(ours.named_parameters()), dict(ref.named_parameters())
    assert list(a) == list(b)
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    # the class is the reference's own (unmodified file), not this repository's re-write
    assert type(ours).__module__ == "models" and type(ours).forward.__code__.co_filename.endswith("oracle/_ref/models.py")
