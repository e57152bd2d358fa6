"""Import the staged reference (``oracle/_ref/``, see stage_ref.py).  TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Two ways to load the reference's ``models.py`` (never edited):
  * ``load("reference")``: over its OWN ``layers/attention.py`` / ``layers/encoding.py`` -- the real reference, CPU or eager CUDA;
  * ``load("b200")``:      over ``mmbidaf_b200.layers`` -- the drop-in contract of north_star ("keep the models.py MMBiDAF module
                           ... unchanged"): ``from layers.encoding import *`` / ``from layers.attention import *`` (models.py:4-5)
                           resolve to this repository's layers.
The frozen ResNet-101 of ``ImageEmbedding`` (layers/encoding.py:124) needs a weight download and is outside the hot path
(SURVEY.md section 8c): the constructor is pointed at ``weights=None`` while the model is built and ``stub_resnet`` replaces the
module by ``Flatten`` -- images are fed as (B, Li, 1000, 1, 1) feature rows, which models.py:105-110 reshapes untouched.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from contextlib import contextmanager

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF, f)) for f in ("models.py", "layers/attention.py", "layers/encoding.py"))


def _load_file(name: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@contextmanager
def _no_resnet_download():
    import torchvision
    original = torchvision.models.resnet101
    torchvision.models.resnet101 = lambda pretrained=True, **kw: original(weights=None)
    try:
        yield
    finally:
        torchvision.models.resnet101 = original


_cache = {}


def load(layers: str = "reference"):
    """-> (models module, layers.attention module, layers.encoding module) with ``layers`` = "reference" | "b200"."""
    if layers in _cache:
        return _cache[layers]
    if not available():
        raise FileNotFoundError(f"{REF} is not staged: run `python oracle/stage_ref.py` where /root/reference is mounted")
    saved = {k: sys.modules.get(k) for k in ("layers", "layers.attention", "layers.encoding", "models")}
    try:
        if layers == "reference":
            pkg = types.ModuleType("layers")
            pkg.__path__ = [os.path.join(REF, "layers")]
            sys.modules["layers"] = pkg
            att = _load_file("layers.attention", os.path.join(REF, "layers", "attention.py"))
            enc = _load_file("layers.encoding", os.path.join(REF, "layers", "encoding.py"))
        elif layers == "b200":
            import mmbidaf_b200.layers as pkg
            import mmbidaf_b200.layers.attention as att
            import mmbidaf_b200.layers.encoding as enc
            sys.modules["layers"], sys.modules["layers.attention"], sys.modules["layers.encoding"] = pkg, att, enc
        else:
            raise ValueError(layers)
        models = _load_file("models", os.path.join(REF, "models.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache[layers] = (models, att, enc)
    return _cache[layers]


def build_model(layers, hidden, e_text, e_audio, e_image, device, drop_prob=0.0, max_transcript_length=409, params=None):
    """The reference's MMBiDAF (models.py:29) over the chosen layers, ResNet stubbed, optionally with a state dict loaded."""
    import torch
    models, _, _ = load(layers)
    with _no_resnet_download():
        model = models.MMBiDAF(hidden, e_text, e_audio, e_image, device, drop_prob=drop_prob,
                               max_transcript_length=max_transcript_length)
    model.image_keyframes_emb = torch.nn.Flatten(1)
    if params is not None:
        missing, unexpected = model.load_state_dict(params, strict=True)
        assert not missing and not unexpected, (missing, unexpected)
    return model.to(device)
