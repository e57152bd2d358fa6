"""north_star's drop-in contract, executed: the reference's OWN ``models.py`` (oracle/_ref/models.py, staged byte for byte by
oracle/stage_ref.py, never edited) runs on the B200 with ``from layers.encoding import *`` / ``from layers.attention import *``
(models.py:4-5) resolving to ``mmbidaf_b200.layers`` -- its forward (models.py:94-206) with CPU-built masks moved to the device
(:116-129), ``forward()`` of the decoder per step (:163, :183), the Python ``int(tensor)`` loss loops (:166-173, :186-193) --
and reproduces the reference's golden outputs: train loss and every gradient, eval distributions, arg-max indices (bit-exact)."""
import pytest
import torch

from conftest import grad_err, load_golden, rel_err
from mmbidaf_b200.synth import Batch, make_batch
from oracle import mmbidaf_oracle as O
from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref is not staged (python oracle/stage_ref.py)")]
TOL = 1e-5


def _run(model, batch, train):
    model.train(train)
    model.zero_grad()
    b = batch.to("cuda")
    return model(b.text, b.text_len, b.audio, b.audio_len, b.images, b.image_len, b.targets, b.target_len, b.max_dec_len)


def test_reference_models_py_over_b200_layers_small_golden():
    g = load_golden("model_small.pt")
    hidden, e_t, e_a, e_i, m = g["dims"]
    model = ref_loader.build_model("b200", hidden, e_t, e_a, e_i, torch.device("cuda"), 0.0, m, params=g["params"])
    assert type(model).forward.__code__.co_filename.endswith("oracle/_ref/models.py")
    batch = Batch(**g["batch"])
    out, loss = _run(model, batch, True)
    assert rel_err(out, g["train_out"]) < TOL and rel_err(loss, g["train_loss"]) < TOL
    loss.backward()
    seen = 0
    for name, p in model.named_parameters():
        if name in g["train_grads"]:
            assert grad_err(p.grad, g["train_grads"][name], name) < 2e-4, name
            seen += 1
    assert seen == len(g["train_grads"])
    with torch.no_grad():
        out_e, loss_e = _run(model, batch, False)
    assert rel_err(out_e, g["eval_out"]) < TOL and rel_err(loss_e, g["eval_loss"]) < TOL
    assert torch.equal(out_e.argmax(dim=2).cpu(), g["eval_argmax"])
    # a second training call on the same module (fresh encoder tensors -> fresh decoder sequence cache): same numbers
    out2, loss2 = _run(model, batch, True)
    assert torch.equal(out2, out) and torch.equal(loss2, loss)


def test_reference_models_py_over_b200_layers_readme_golden():
    g = load_golden("model_readme.pt")
    hidden, e_t, e_a, e_i, m = g["dims"]
    params = O.make_params(hidden, e_t, e_a, e_i, m, seed=g["param_seed"])
    model = ref_loader.build_model("b200", hidden, e_t, e_a, e_i, torch.device("cuda"), 0.0, m, params=params)
    batch = make_batch(*g["batch_shape"], e_t, e_a, e_i, seed=g["batch_seed"])
    out, loss = _run(model, batch, True)
    assert rel_err(out, g["train_out"]) < TOL and rel_err(loss, g["train_loss"]) < TOL
    loss.backward()
    grads = dict(model.named_parameters())
    for name, want in g["train_grads_sample"].items():
        assert grad_err(grads[name].grad, want, name) < 2e-4, name
    for name, want in g["train_grad_norms"].items():
        got = float(grads[name].grad.double().norm())
        assert abs(got - want) <= 2e-3 * max(want, 1e-3), name
    with torch.no_grad():
        out_e, loss_e = _run(model, batch, False)
    assert rel_err(out_e, g["eval_out"]) < TOL and rel_err(loss_e, g["eval_loss"]) < TOL
    assert torch.equal(out_e.argmax(dim=2).cpu(), g["eval_argmax"])


def test_reference_models_py_matches_own_models_py_on_both_tiers():
    """Same weights: the reference's models.py over our layers and this repository's re-written models.py (device-side masks and
    gathers, fused loss terms, side streams) agree at README sizes on both precision tiers.  (Dropout off: the two forward
    passes visit the encoders in a different order, so they would draw different masks from the generator.)"""
    import mmbidaf_b200
    from mmbidaf_b200.models import MMBiDAF
    dims = (100, 300, 128, 1000, 409)
    params = O.make_params(*dims, seed=3)
    batch = make_batch(4, 37, 70, 11, 5, seed=4)
    ref_model = ref_loader.build_model("b200", *dims[:4], torch.device("cuda"), 0.0, dims[4], params=params)
    own = MMBiDAF(*dims[:4], torch.device("cuda"), drop_prob=0.0, max_transcript_length=dims[4])
    own.load_state_dict(params)
    own = own.cuda()
    for tier, tol in (("fp32", 1e-5), ("fast", 2e-2)):
        mmbidaf_b200.set_precision(tier)
        try:
            out_r, loss_r = _run(ref_model, batch, True)
            loss_r.backward()
            out_o, loss_o = _run(own, batch, True)
            loss_o.backward()
        finally:
            mmbidaf_b200.set_precision("fp32")
        assert rel_err(out_r, out_o) < tol and rel_err(loss_r, loss_o) < tol, tier
        go = dict(own.named_parameters())
        for name, p in ref_model.named_parameters():
            if p.grad is not None and name.endswith(("text_weight", "weight_hh_l0", "out.weight", "proj.weight")):
                assert grad_err(p.grad, go[name].grad, name) < max(tol * 5, 2e-4), (tier, name)
