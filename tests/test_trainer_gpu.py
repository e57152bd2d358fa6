"""GPU: the training step (both precision tiers) and its CUDA-graph replay."""
import pytest
import torch

from mmbidaf_b200.synth import make_batch

pytestmark = pytest.mark.gpu
DIMS = (100, 300, 128, 1000, 64)


def _make(seed=3, drop=0.0):
    from mmbidaf_b200.models import MMBiDAF
    torch.manual_seed(seed)
    return MMBiDAF(*DIMS[:4], torch.device("cuda"), drop_prob=drop, max_transcript_length=DIMS[4]).cuda()


@pytest.mark.parametrize("tier", ["fp32", "fast"])
def test_graphed_step_is_bit_identical_to_eager(tier):
    import mmbidaf_b200
    from mmbidaf_b200.trainer import Trainer
    mmbidaf_b200.set_precision(tier)
    try:
        batch = make_batch(5, 33, 70, 9, 4, seed=11).to("cuda")
        work = torch.cuda.Stream()
        with torch.cuda.stream(work):
            eager, graphed = Trainer(_make()), Trainer(_make())
            for _ in range(5):
                loss_e = eager.step(batch)
            graphed.capture(batch, warmup=2)
            for _ in range(3):
                loss_g = graphed.step_graphed()
            torch.cuda.synchronize()
        assert torch.isfinite(loss_e) and torch.equal(loss_e, loss_g)
        for p, q in zip(eager.model.parameters(), graphed.model.parameters()):
            assert torch.equal(p, q)
        with pytest.raises(ValueError, match="lengths differ"):
            graphed.step_graphed(make_batch(5, 33, 70, 9, 4, seed=12).to("cuda"))
    finally:
        mmbidaf_b200.set_precision("fp32")


def test_training_reduces_the_loss_and_keeps_parameters_aligned():
    from mmbidaf_b200.trainer import Trainer
    batch = make_batch(4, 20, 30, 6, 3, seed=5).to("cuda")
    tr = Trainer(_make(drop=0.2))
    assert all(p.data_ptr() % 16 == 0 for p in tr.grads.params)       # kernels read weights with 128-bit loads
    first = float(tr.step(batch))
    for _ in range(15):
        last = float(tr.step(batch))
    assert last < first


@pytest.mark.parametrize("tier", ["fp32", "fast"])
def test_leaf_lanes_give_the_same_gradients_as_the_inline_backward(tier):
    """Inside ``functional.leaf_lanes()`` the weight-gradient products run on side streams (same kernels, same operands):
    every gradient must be bit-identical to the plain single-stream backward, run after run."""
    import contextlib
    import mmbidaf_b200
    from mmbidaf_b200 import functional as Fn
    mmbidaf_b200.set_precision(tier)
    try:
        model = _make(drop=0.0)
        model.train()
        params = [p for p in model.parameters() if p.requires_grad]
        b = make_batch(6, 40, 90, 11, 5, seed=21).to("cuda")

        def grads(lanes):
            _, loss = model(b.text, b.text_len, b.audio, b.audio_len, b.images, b.image_len, b.targets, b.target_len, b.max_dec_len)
            with (Fn.leaf_lanes() if lanes else contextlib.nullcontext()):
                g = torch.autograd.grad(loss, params, allow_unused=True)
            out = [None if t is None else t.clone() for t in g]
            torch.cuda.synchronize()
            return out

        want = grads(False)
        assert sum(t is not None for t in want) > 40
        for _ in range(4):
            got = grads(True)
            for w, g in zip(want, got):
                assert (w is None) == (g is None)
                if w is not None:
                    assert torch.equal(w, g)
    finally:
        mmbidaf_b200.set_precision("fp32")
