"""GPU: BASELINE config 5 (long-lecture evaluate.py-style extractive decoding): parity at a reduced long shape,
size-independent properties at the full shape (B=16, Lt=4096, La=4096, Li=2048, M=4096, 16 greedy steps)."""
import pytest
import torch

from conftest import rel_err
from mmbidaf_b200.synth import make_batch
from oracle import mmbidaf_oracle as O

pytestmark = pytest.mark.gpu


def _eval(model, batch):
    model.eval()
    b = batch.to("cuda")
    with torch.no_grad():
        return model(b.text, b.text_len, b.audio, b.audio_len, b.images, b.image_len, b.targets, b.target_len, b.max_dec_len)


def _model(m, seed=224):
    from mmbidaf_b200.models import MMBiDAF
    params = O.make_params(100, 300, 128, 1000, m, seed=seed)
    model = MMBiDAF(100, 300, 128, 1000, torch.device("cuda"), drop_prob=0.0, max_transcript_length=m)
    model.load_state_dict(params)
    return model.cuda(), params


def test_reduced_long_shape_matches_oracle():
    m = 700
    model, params = _model(m)
    batch = make_batch(2, 700, 900, 300, 6, seed=31)
    out, loss = _eval(model, batch)
    with torch.no_grad():
        want, want_loss = O.mmbidaf_forward(params, batch.text, batch.text_len, batch.audio, batch.audio_len,
                                            batch.images.flatten(2), batch.image_len, batch.targets, batch.max_dec_len, m,
                                            training=False, fast_lstm=True)
    assert rel_err(out, want) < 1e-4 and rel_err(loss, want_loss) < 1e-4       # fp32 tier, 900-step recurrences
    assert torch.equal(out.argmax(dim=2).cpu(), want.argmax(dim=2))             # selected sentences: bit-exact


@pytest.mark.parametrize("tier", ["fp32", "fast"])
def test_full_config5_properties(tier):
    import mmbidaf_b200
    from mmbidaf_b200.decode import get_generated_indices
    mmbidaf_b200.set_precision(tier)
    try:
        m = 4096
        model, _ = _model(m)
        batch = make_batch(16, 4096, 4096, 2048, 16, seed=32)
        out, loss = _eval(model, batch)
        torch.cuda.synchronize()
    finally:
        mmbidaf_b200.set_precision("fp32")
    assert out.shape == (16, 16, m) and torch.isfinite(out).all() and torch.isfinite(loss)
    assert (out.sum(dim=2) - 1).abs().max() < 1e-4                              # every step is a distribution ...
    lens = torch.tensor(batch.text_len, device="cuda")
    beyond = torch.arange(m, device="cuda").view(1, 1, m) >= lens.view(-1, 1, 1)
    assert (out.masked_select(beyond.expand_as(out)) == 0).all()                # ... with exactly zero mass past each length
    picks = get_generated_indices(out, batch.text_len)
    assert all(0 <= k < n - 1 for ks, n in zip(picks, batch.text_len) for k in ks)
