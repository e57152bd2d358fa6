"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only in the build container, where the reference tree is mounted read-only at
/root/reference.  It imports the reference's own ``layers.attention``,
``layers.encoding`` and ``models`` modules (unmodified), feeds them seeded inputs and
stores inputs + outputs as small ``.pt`` files.  Nothing at test time reads
/root/reference: the tests replay these files.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt

The frozen ResNet-101 (layers/encoding.py:124) needs a weight download; it is outside
the hot path, so the constructor is pointed at ``weights=None`` and the module is
replaced by ``Flatten`` (images are fed as (B, Li, E, 1, 1) feature rows), exactly as
SURVEY.md section 8c describes.  Generated with torch 2.11.0 (CPU, fp32).
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MMBIDAF_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from mmbidaf_b200.synth import make_batch  # noqa: E402
from oracle.mmbidaf_oracle import make_params  # noqa: E402


def import_reference():
    import torchvision
    original = torchvision.models.resnet101
    torchvision.models.resnet101 = lambda pretrained=True: original(weights=None)
    sys.path.insert(0, REF)
    for name in ("layers", "layers.attention", "layers.encoding", "models"):
        sys.modules.pop(name, None)
    import layers.attention as ref_att
    import layers.encoding as ref_enc
    import models as ref_models
    sys.path.remove(REF)
    return ref_att, ref_enc, ref_models


def save(name, blob):
    path = os.path.join(HERE, name)
    torch.save(blob, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def clone_state(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def masks_from(lengths, max_len):
    return torch.arange(max_len).unsqueeze(0) < torch.tensor(lengths).unsqueeze(1)


def bidaf_cases(ref_att):
    for tag, (bsz, lc, lq, d, c_len, q_len) in {
        "small": (3, 7, 5, 8, [7, 4, 1], [5, 2, 3]),
        "d200": (2, 33, 17, 200, [33, 20], [17, 9]),
    }.items():
        torch.manual_seed(11)
        mod = ref_att.BiDAFAttention(d, drop_prob=0.2)
        with torch.no_grad():
            mod.bias.fill_(0.37)
        text = torch.randn(bsz, lc, d, requires_grad=True)
        modality = torch.randn(bsz, lq, d, requires_grad=True)
        c_mask, q_mask = masks_from(c_len, lc), masks_from(q_len, lq)
        mod.eval()
        out = mod(text, modality, c_mask, q_mask)
        sim = mod.get_similarity_matrix(text, modality)
        grad_out = torch.randn_like(out)
        out.backward(grad_out)
        blob = {"state": clone_state(mod), "text": text.detach(), "modality": modality.detach(),
                "text_mask": c_mask, "modality_mask": q_mask, "out": out.detach(), "similarity": sim.detach(),
                "grad_out": grad_out, "grad_text": text.grad.clone(), "grad_modality": modality.grad.clone(),
                "grad_params": {k: v.grad.clone() for k, v in mod.named_parameters()}}
        # training mode: dropout inside get_similarity_matrix only (attention.py:66-67)
        mod.train()
        mod.zero_grad()
        text.grad = None
        modality.grad = None
        torch.manual_seed(5)
        out_t = mod(text, modality, c_mask, q_mask)
        out_t.backward(grad_out)
        blob.update({"train_seed": 5, "train_drop_prob": 0.2, "train_out": out_t.detach(),
                     "train_grad_text": text.grad.clone(), "train_grad_modality": modality.grad.clone(),
                     "train_grad_params": {k: v.grad.clone() for k, v in mod.named_parameters()}})
        save(f"bidaf_{tag}.pt", blob)

    # masked softmax on its own, both directions and the (B, M) decoder use
    torch.manual_seed(3)
    logits = torch.randn(2, 4, 6)
    row_mask = masks_from([6, 3], 6).unsqueeze(1)
    col_mask = masks_from([4, 1], 4).unsqueeze(2)
    flat = torch.randn(3, 9)
    flat_mask = masks_from([9, 4, 0], 9)            # an all-masked row -> uniform 1/9
    save("masked_softmax.pt", {
        "logits": logits, "row_mask": row_mask, "col_mask": col_mask,
        "row": ref_att.masked_softmax(logits, row_mask, dim=2), "col": ref_att.masked_softmax(logits, col_mask, dim=1),
        "flat": flat, "flat_mask": flat_mask, "flat_out": ref_att.masked_softmax(flat, flat_mask),
        "flat_log": ref_att.masked_softmax(flat, flat_mask, log_softmax=True)})


def rnn_cases(ref_enc):
    for tag, (in_size, hid, layers, lengths, max_len) in {
        "l1": (6, 5, 1, [4, 7, 4, 7, 2], 7),
        "l2": (12, 5, 2, [3, 9, 9, 1, 5, 3], 9),
        "h100": (100, 100, 1, [13, 20, 7], 20),
    }.items():
        torch.manual_seed(21)
        mod = ref_enc.RNNEncoder(in_size, hid, layers, drop_prob=0.0)
        mod.eval()
        x = torch.randn(len(lengths), max_len, in_size)
        for b, n in enumerate(lengths):
            x[b, n:] = 0
        x.requires_grad_(True)
        out, h_n = mod(x, lengths)
        g_out, g_h = torch.randn_like(out), torch.randn_like(h_n)
        (out * g_out).sum().add((h_n * g_h).sum()).backward()
        save(f"rnn_{tag}.pt", {"state": clone_state(mod), "x": x.detach(), "lengths": lengths, "layers": layers,
                               "out": out.detach(), "h_n": h_n.detach(), "grad_out": g_out, "grad_h_n": g_h,
                               "grad_x": x.grad.clone(),
                               "grad_params": {k: v.grad.clone() for k, v in mod.named_parameters()}})


def embedding_case(ref_enc):
    torch.manual_seed(31)
    mod = ref_enc.Embedding(embedding_size=10, hidden_size=6, drop_prob=0.0)
    mod.eval()
    x = torch.randn(2, 5, 10)
    save("embedding.pt", {"state": clone_state(mod), "x": x, "out": mod(x).detach()})


def decoder_case(ref_att):
    torch.manual_seed(41)
    e, h, m, lt, bsz = 7, 5, 9, 6, 3
    mod = ref_att.MultimodalAttentionDecoder(e, h, m, num_layers=1)
    mod.eval()
    enc_a = torch.randn(bsz, lt, 2 * h, requires_grad=True)
    enc_i = torch.randn(bsz, lt, 2 * h, requires_grad=True)
    hid = torch.randn(bsz, 1, h, requires_grad=True)
    cell = torch.zeros(1, bsz, h)
    cov = torch.zeros(bsz, lt, 1)
    mask = torch.cat([masks_from([6, 3, 5], lt), torch.zeros(bsz, m - lt, dtype=torch.bool)], dim=1)
    sent = [torch.randn(bsz, 1, e), torch.randn(bsz, 1, e)]
    steps = []
    state = (hid, cell, cov)
    total = 0
    for k in range(2):
        probs, h1, c1, att, cov1 = mod(sent[k], state[0], state[1], enc_a, enc_i, state[2], mask)
        steps.append({"probs": probs.detach(), "h": h1.detach(), "cell": c1.detach(), "att_cov": att.detach(),
                      "coverage": cov1.detach()})
        total = total - torch.log(probs[:, k] + 1e-12).sum() + torch.min(att, cov1).sum()
        state = (h1, c1, cov1)
    total.backward()
    save("decoder_small.pt", {"state": clone_state(mod), "enc_a": enc_a.detach(), "enc_i": enc_i.detach(),
                              "h0": hid.detach(), "cell0": cell, "cov0": cov, "mask": mask, "sent": sent,
                              "steps": steps, "loss": total.detach(), "grad_enc_a": enc_a.grad.clone(),
                              "grad_enc_i": enc_i.grad.clone(), "grad_h0": hid.grad.clone(),
                              "grad_params": {k: v.grad.clone() for k, v in mod.named_parameters()}})


def build_reference_model(ref_models, hidden, e_text, e_audio, e_image, m, seed):
    model = ref_models.MMBiDAF(hidden, e_text, e_audio, e_image, torch.device("cpu"), drop_prob=0.0,
                               max_transcript_length=m)
    model.image_keyframes_emb = torch.nn.Flatten(1)
    params = make_params(hidden, e_text, e_audio, e_image, m, seed=seed)
    missing, unexpected = model.load_state_dict(params, strict=True)
    assert not missing and not unexpected
    return model, params


def run_model(model, batch, train):
    model.train(train)
    model.zero_grad()
    out, loss = model(batch.text, batch.text_len, batch.audio, batch.audio_len, batch.images, batch.image_len,
                      batch.targets, batch.target_len, batch.max_dec_len)
    grads = None
    if train:
        loss.backward()
        grads = {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}
    return out.detach(), loss.detach(), grads


def model_cases(ref_models):
    # tiny model: everything stored
    hidden, e_t, e_a, e_i, m = 6, 10, 4, 12, 11
    model, params = build_reference_model(ref_models, hidden, e_t, e_a, e_i, m, seed=51)
    batch = make_batch(3, 8, 9, 5, 4, e_t, e_a, e_i, seed=52)
    out_t, loss_t, grads = run_model(model, batch, True)
    with torch.no_grad():
        out_e, loss_e, _ = run_model(model, batch, False)
    save("model_small.pt", {"dims": (hidden, e_t, e_a, e_i, m), "param_seed": 51, "batch_seed": 52,
                            "batch_shape": (3, 8, 9, 5, 4), "params": params, "batch": batch.__dict__,
                            "train_out": out_t, "train_loss": loss_t, "train_grads": grads,
                            "eval_out": out_e, "eval_loss": loss_e, "eval_argmax": out_e.argmax(dim=2)})

    # README sizes (README.md:38-44): params/batch regenerated from seeds, outputs + grad norms stored
    hidden, e_t, e_a, e_i, m = 100, 300, 128, 1000, 409
    model, params = build_reference_model(ref_models, hidden, e_t, e_a, e_i, m, seed=224)
    batch = make_batch(3, 24, 40, 9, 5, e_t, e_a, e_i, seed=225)
    out_t, loss_t, grads = run_model(model, batch, True)
    with torch.no_grad():
        out_e, loss_e, _ = run_model(model, batch, False)
    keep = ("bidaf_att_audio.text_weight", "bidaf_att_image.text_modality_weight", "multimodal_att_decoder.v1.weight",
            "text_enc.rnn.bias_hh_l0_reverse", "mod_t_a.rnn.bias_ih_l1", "emb.hwy.gates.0.bias")
    save("model_readme.pt", {"dims": (hidden, e_t, e_a, e_i, m), "param_seed": 224, "batch_seed": 225,
                             "batch_shape": (3, 24, 40, 9, 5),
                             "param_checksum": float(sum(v.double().sum() for v in params.values())),
                             "train_out": out_t, "train_loss": loss_t,
                             "train_grad_norms": {k: float(v.double().norm()) for k, v in grads.items()},
                             "train_grads_sample": {k: grads[k] for k in keep},
                             "eval_out": out_e, "eval_loss": loss_e, "eval_argmax": out_e.argmax(dim=2)})


def main():
    torch.set_num_threads(1)          # deterministic summation order for the stored values
    ref_att, ref_enc, ref_models = import_reference()
    bidaf_cases(ref_att)
    rnn_cases(ref_enc)
    embedding_case(ref_enc)
    decoder_case(ref_att)
    model_cases(ref_models)
    save("MANIFEST.pt", {"torch": torch.__version__, "reference": REF})


if __name__ == "__main__":
    main()
