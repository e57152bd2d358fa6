"""The oracle (oracle/mmbidaf_oracle.py) against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import torch
import torch.nn.functional as F

from conftest import grad_err, load_golden, rel_err
from mmbidaf_b200.synth import make_batch
from oracle import mmbidaf_oracle as O

TOL = 2e-6     # same library, same dtype, different op grouping


def test_masked_softmax_matches_reference():
    g = load_golden("masked_softmax.pt")
    assert torch.equal(O.masked_softmax(g["logits"], g["row_mask"], dim=2), g["row"])
    assert torch.equal(O.masked_softmax(g["logits"], g["col_mask"], dim=1), g["col"])
    assert torch.equal(O.masked_softmax(g["flat"], g["flat_mask"]), g["flat_out"])
    assert torch.equal(O.masked_softmax(g["flat"], g["flat_mask"], log_softmax=True), g["flat_log"])
    # known answers: exact zeros at masked slots, uniform when everything is masked
    assert (g["flat_out"][1, 4:] == 0).all()
    assert torch.allclose(g["flat_out"][2], torch.full((9,), 1 / 9))


def _bidaf_grads(g, **kw):
    p = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
    text = g["text"].clone().requires_grad_(True)
    modality = g["modality"].clone().requires_grad_(True)
    out = O.bidaf_attention(p, text, modality, g["text_mask"], g["modality_mask"], **kw)
    out.backward(g["grad_out"])
    return out.detach(), text.grad, modality.grad, {k: v.grad for k, v in p.items()}


def test_bidaf_eval_matches_reference():
    for name in ("bidaf_small.pt", "bidaf_d200.pt"):
        g = load_golden(name)
        assert rel_err(O.bidaf_similarity(g["state"], g["text"], g["modality"]), g["similarity"]) < TOL
        out, gt, gm, gp = _bidaf_grads(g)
        assert rel_err(out, g["out"]) < TOL
        assert rel_err(gt, g["grad_text"]) < 1e-5 and rel_err(gm, g["grad_modality"]) < 1e-5
        for k, v in g["grad_params"].items():
            assert grad_err(gp[k], v, k) < 1e-5, k      # bias grad is identically 0: floor handles it
        # reassociated b = s1 (s2^T c) is the same function
        out_r, *_ = _bidaf_grads(g, reassociate=True)
        assert rel_err(out_r, g["out"]) < 1e-5


def test_bidaf_padded_text_rows_keep_attention_block():
    g = load_golden("bidaf_small.pt")
    d = g["text"].shape[2]
    pad_rows = g["out"][1, 4:]                      # sample 1 has text length 4
    assert pad_rows[:, d:2 * d].abs().max() > 0     # 'a' block is a full row-softmax even on padded rows
    # column soft-max gives padded rows zero weight, so b does not depend on them
    text2 = g["text"].clone()
    text2[1, 4:] += 1.0
    a = O.bidaf_attention(g["state"], g["text"], g["modality"], g["text_mask"], g["modality_mask"])
    b = O.bidaf_attention(g["state"], text2, g["modality"], g["text_mask"], g["modality_mask"])
    assert torch.allclose(a[1, :4], b[1, :4], atol=1e-6)


def test_bidaf_training_dropout_order_matches_reference():
    """attention.py:66-67 draws the text mask first, then the modality mask; F.dropout of a
    ones tensor under the same seed reproduces both keep-masks."""
    for name in ("bidaf_small.pt", "bidaf_d200.pt"):
        g = load_golden(name)
        pr = g["train_drop_prob"]
        torch.manual_seed(g["train_seed"])
        keep_c = (F.dropout(torch.ones_like(g["text"]), pr, True) != 0).float()
        keep_q = (F.dropout(torch.ones_like(g["modality"]), pr, True) != 0).float()
        out, gt, gm, gp = _bidaf_grads(g, keep_text=keep_c, keep_modality=keep_q, drop_prob=pr)
        assert rel_err(out, g["train_out"]) < TOL
        assert rel_err(gt, g["train_grad_text"]) < 1e-5 and rel_err(gm, g["train_grad_modality"]) < 1e-5
        for k, v in g["train_grad_params"].items():
            assert grad_err(gp[k], v, k) < 1e-5, k


def test_embedding_matches_reference():
    g = load_golden("embedding.pt")
    assert rel_err(O.embedding(g["state"], g["x"]), g["out"]) < TOL


def test_rnn_encoder_matches_reference():
    for name in ("rnn_l1.pt", "rnn_l2.pt", "rnn_h100.pt"):
        g = load_golden(name)
        p = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
        x = g["x"].clone().requires_grad_(True)
        out, h_n = O.rnn_encoder(p, x, g["lengths"], g["layers"])
        assert rel_err(out, g["out"]) < 1e-5, name
        assert rel_err(h_n, g["h_n"]) < 1e-5, name           # sorted-order rows (Q3)
        for b, n in enumerate(g["lengths"]):
            assert (out[b, n:] == 0).all()                   # exact zeros past the length
        ((out * g["grad_out"]).sum() + (h_n * g["grad_h_n"]).sum()).backward()
        assert rel_err(x.grad, g["grad_x"]) < 1e-4, name
        for k, v in g["grad_params"].items():
            assert rel_err(p[k].grad, v) < 1e-4, (name, k)
        out2, h2 = O.rnn_encoder_aten(g["state"], g["x"], g["lengths"], g["layers"])
        assert rel_err(out2, g["out"]) < TOL and rel_err(h2, g["h_n"]) < TOL


def test_decoder_steps_match_reference():
    g = load_golden("decoder_small.pt")
    p = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
    enc_a = g["enc_a"].clone().requires_grad_(True)
    enc_i = g["enc_i"].clone().requires_grad_(True)
    h = g["h0"].clone().requires_grad_(True)
    state = (h, g["cell0"], g["cov0"])
    loss = 0
    for k, want in enumerate(g["steps"]):
        probs, h1, c1, att, cov = O.decoder_step(p, g["sent"][k], state[0], state[1], enc_a, enc_i, state[2], g["mask"])
        for got, key in ((probs, "probs"), (h1, "h"), (c1, "cell"), (att, "att_cov"), (cov, "coverage")):
            assert got.shape == want[key].shape
            assert rel_err(got, want[key]) < TOL, (k, key)
        assert (probs[~g["mask"]] == 0).all()
        loss = loss - torch.log(probs[:, k] + 1e-12).sum() + torch.min(att, cov).sum()
        state = (h1, c1, cov)
    assert rel_err(loss, g["loss"]) < TOL
    loss.backward()
    assert rel_err(enc_a.grad, g["grad_enc_a"]) < 1e-5 and rel_err(enc_i.grad, g["grad_enc_i"]) < 1e-5
    assert rel_err(h.grad, g["grad_h0"]) < 1e-5
    for k, v in g["grad_params"].items():
        assert grad_err(p[k].grad, v, k) < 1e-5, k


def _batch_from(g):
    from mmbidaf_b200.synth import Batch
    return Batch(**g["batch"])


def test_model_small_matches_reference():
    g = load_golden("model_small.pt")
    hidden, e_t, e_a, e_i, m = g["dims"]
    params = O.make_params(hidden, e_t, e_a, e_i, m, seed=g["param_seed"])
    for k, v in g["params"].items():
        assert torch.equal(params[k], v), k               # seed -> parameters is reproducible
    b = _batch_from(g)
    regen = make_batch(*g["batch_shape"], e_t, e_a, e_i, seed=g["batch_seed"])
    assert torch.equal(regen.text, b.text) and regen.text_len == b.text_len and torch.equal(regen.targets, b.targets)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    img = b.images.flatten(2)
    out, loss = O.mmbidaf_forward(p, b.text, b.text_len, b.audio, b.audio_len, img, b.image_len, b.targets,
                                  b.max_dec_len, m, training=True)
    assert rel_err(out, g["train_out"]) < 1e-5 and rel_err(loss, g["train_loss"]) < 1e-5
    loss.backward()
    for k, v in g["train_grads"].items():
        assert grad_err(p[k].grad, v, k) < 2e-4, k
    with torch.no_grad():
        for fast in (False, True):
            out_e, loss_e = O.mmbidaf_forward(params, b.text, b.text_len, b.audio, b.audio_len, img, b.image_len,
                                              b.targets, b.max_dec_len, m, training=False, fast_lstm=fast)
            assert rel_err(out_e, g["eval_out"]) < 1e-5 and rel_err(loss_e, g["eval_loss"]) < 1e-5
            assert torch.equal(out_e.argmax(dim=2), g["eval_argmax"])        # selected sentence indices: bit-exact


def test_model_readme_sizes_match_reference():
    g = load_golden("model_readme.pt")
    hidden, e_t, e_a, e_i, m = g["dims"]
    params = O.make_params(hidden, e_t, e_a, e_i, m, seed=g["param_seed"])
    assert abs(float(sum(v.double().sum() for v in params.values())) - g["param_checksum"]) < 1e-9
    b = make_batch(*g["batch_shape"], e_t, e_a, e_i, seed=g["batch_seed"])
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    img = b.images.flatten(2)
    out, loss = O.mmbidaf_forward(p, b.text, b.text_len, b.audio, b.audio_len, img, b.image_len, b.targets,
                                  b.max_dec_len, m, training=True, fast_lstm=True)
    assert rel_err(out, g["train_out"]) < 1e-5 and rel_err(loss, g["train_loss"]) < 1e-5
    loss.backward()
    for k, v in g["train_grads_sample"].items():
        assert grad_err(p[k].grad, v, k) < 2e-4, k
    for k, v in g["train_grad_norms"].items():
        assert abs(float(p[k].grad.double().norm()) - v) <= 2e-3 * max(v, 1e-3), k
    with torch.no_grad():
        out_e, loss_e = O.mmbidaf_forward(params, b.text, b.text_len, b.audio, b.audio_len, img, b.image_len,
                                          b.targets, b.max_dec_len, m, training=False, fast_lstm=True)
    assert rel_err(out_e, g["eval_out"]) < 1e-5 and rel_err(loss_e, g["eval_loss"]) < 1e-5
    assert torch.equal(out_e.argmax(dim=2), g["eval_argmax"])


def test_greedy_indices_stop_at_eos():
    dist = torch.zeros(4, 6)
    dist[0, 2] = 1
    dist[1, 0] = 1
    dist[2, 4] = 1          # text_len 5 -> EOS row 4
    dist[3, 1] = 1
    assert O.greedy_indices(dist, 5) == [2, 0]
    assert O.greedy_indices(dist, 6) == [2, 0, 4, 1]


def test_greedy_search_matches_the_reference_evaluate_py():
    """Selected sentence indices pinned to the reference's own evaluate.py:167-202, :236-259 (tests/golden/make_golden_greedy.py
    imports it unmodified and runs it against pickled transcripts): EOS at the first step, an index beyond the transcript
    (skipped, not a stop), the same sentence picked repeatedly."""
    from mmbidaf_b200.decode import get_generated_indices, greedy_search
    g = load_golden("greedy_search.pt")
    assert len(g["cases"]) == 3
    for case in g["cases"]:
        dist, lengths, want = case["dist"], case["lengths"], case["indices"]
        assert [O.greedy_indices(dist[b], lengths[b]) for b in range(len(lengths))] == want
        assert [greedy_search(dist[b], lengths[b]) for b in range(len(lengths))] == want
        assert get_generated_indices(dist, lengths) == want
    assert any(len(w) == 0 for c in g["cases"] for w in c["indices"])
