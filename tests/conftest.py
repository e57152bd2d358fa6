import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(got: torch.Tensor, want: torch.Tensor) -> float:
    """max |got - want| / max(|want|, tiny): the 'rel' of north_star's tolerances."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    denom = max(float(want.abs().max()), 1e-30)
    return float((got - want).abs().max()) / denom


NULL_GRADS = ("bidaf_att_audio.bias", "bidaf_att_image.bias", "v1.bias", "v2.bias")


def grad_err(got: torch.Tensor, want: torch.Tensor, name: str = "", floor: float = 1e-3) -> float:
    """rel_err with the denominator floored.  Three gradients are identically zero by softmax
    shift invariance (BiDAF ``bias``, decoder ``v1.bias`` / ``v2.bias``) and hold only rounding
    noise in the reference itself: for those the check is |got| small, reported as 0."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    if name == "bias" or name.endswith(NULL_GRADS):
        assert float(got.abs().max()) < 1e-2 * max(1.0, float(want.abs().max()) * 1e3), (name, got, want)
        return 0.0
    return float((got - want).abs().max()) / max(float(want.abs().max()), floor)
