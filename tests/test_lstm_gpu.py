"""GPU parity: persistent bi-LSTM kernels vs the reference's golden vectors and the oracle."""
import pytest
import torch

from conftest import grad_err, load_golden, rel_err
from oracle import mmbidaf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _weights(state, layer, dev):
    ws = []
    for suffix in ("", "_reverse"):
        for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            ws.append(state[f"rnn.{kind}_l{layer}{suffix}"].to(dev).requires_grad_(True))
    return ws


def _run_encoder(state, x, lengths, layers, grad_out=None, grad_h=None):
    """Layer stack exactly as RNNEncoder wires it, straight on the functional op."""
    from mmbidaf_b200 import functional as Fn
    dev = "cuda"
    order = O.sort_order(lengths)
    len_d = torch.tensor(lengths, dtype=torch.int32, device=dev)
    ord_d = order.to(torch.int32).to(dev)
    xin = x.to(dev).requires_grad_(True)
    cur, finals, all_w = xin, [], []
    for k in range(layers):
        w = _weights(state, k, dev)
        all_w.append(w)
        cur, h_n = Fn.lstm_layer(cur, len_d, ord_d, w)
        finals.append(h_n)
    h_all = torch.cat(finals, dim=1)[order.to(dev)]          # Q3: rows left in sorted order
    if grad_out is not None:
        ((cur * grad_out.to(dev)).sum() + (h_all * grad_h.to(dev)).sum()).backward()
    return cur, h_all, xin, all_w


@pytest.mark.parametrize("name", ["rnn_l1.pt", "rnn_l2.pt", "rnn_h100.pt"])
def test_encoder_matches_reference_golden(name):
    g = load_golden(name)
    out, h_n, xin, all_w = _run_encoder(g["state"], g["x"], g["lengths"], g["layers"], g["grad_out"], g["grad_h_n"])
    assert rel_err(out, g["out"]) < TOL and rel_err(h_n, g["h_n"]) < TOL
    for b, n in enumerate(g["lengths"]):
        assert (out[b, n:] == 0).all()                        # pad_packed_sequence zeros, exact
    assert grad_err(xin.grad, g["grad_x"]) < 1e-4
    names = [f"rnn.{kind}_l{{k}}{suffix}" for suffix in ("", "_reverse")
             for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    for k, ws in enumerate(all_w):
        for nm, w in zip(names, ws):
            assert grad_err(w.grad, g["grad_params"][nm.format(k=k)], nm) < 1e-4, (k, nm)


@pytest.mark.parametrize("pair", ["0", "1"])
@pytest.mark.parametrize("cfg", [(1, 1, 3, 4), (7, 33, 20, 16), (5, 50, 100, 100), (40, 19, 100, 100),
                                 (90, 12, 64, 100), (3, 300, 800, 100), (4, 9, 104, 104), (2, 6, 10, 60), (3, 21, 16, 33)])
def test_encoder_matches_oracle(cfg, pair, monkeypatch):
    """pair = 1: the forward recurrence on a two-CTA cluster per (sequence, direction) (bilstm_fwd_pair_kernel; hidden sizes >= 32, odd
    sizes split 17 / 16), pair = 0: one CTA."""
    monkeypatch.setenv("MMB_LSTM_PAIR", pair)
    bsz, max_len, fan_in, hid = cfg
    gen = torch.Generator().manual_seed(31 * bsz + max_len)
    lengths = torch.randint(1, max_len + 1, (bsz,), generator=gen).tolist()
    lengths[0] = max_len
    state = {}
    for suffix in ("", "_reverse"):
        state[f"rnn.weight_ih_l0{suffix}"] = (torch.rand(4 * hid, fan_in, generator=gen) - 0.5) * 0.2
        state[f"rnn.weight_hh_l0{suffix}"] = (torch.rand(4 * hid, hid, generator=gen) - 0.5) * 0.2
        state[f"rnn.bias_ih_l0{suffix}"] = (torch.rand(4 * hid, generator=gen) - 0.5) * 0.2
        state[f"rnn.bias_hh_l0{suffix}"] = (torch.rand(4 * hid, generator=gen) - 0.5) * 0.2
    x = torch.randn(bsz, max_len, fan_in, generator=gen)
    g_out = torch.randn(bsz, max_len, 2 * hid, generator=gen)
    g_h = torch.randn(bsz, 2, hid, generator=gen)
    # oracle in fp64 through the library LSTM (fast) -- itself checked against the explicit loop on CPU
    p64 = {k: v.double().requires_grad_(True) for k, v in state.items()}
    x64 = x.double().requires_grad_(True)
    want_out, want_h = O.rnn_encoder_aten(p64, x64, lengths, 1)
    ((want_out * g_out.double()).sum() + (want_h * g_h.double()).sum()).backward()
    out, h_n, xin, all_w = _run_encoder(state, x, lengths, 1, g_out, g_h)
    assert rel_err(out, want_out) < TOL and rel_err(h_n, want_h) < TOL
    assert grad_err(xin.grad, x64.grad) < 5e-5
    names = [f"rnn.{kind}_l0{suffix}" for suffix in ("", "_reverse")
             for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    for nm, w in zip(names, all_w[0]):
        assert grad_err(w.grad, p64[nm].grad) < 5e-5, nm


@pytest.mark.parametrize("cfg", [(7, 33, 20, 16, 0.2), (5, 50, 100, 100, 0.2), (40, 19, 100, 100, 0.5), (3, 21, 16, 33, 0.35)])
def test_in_kernel_dropout_given_the_mask(cfg):
    """encoding.py:104 `F.dropout(x, drop_prob, training)` (and nn.LSTM's inter-layer dropout) applied INSIDE the recurrence kernels:
    with the same key the dropped output equals out * mask / keep_prob BIT FOR BIT (mask from mmb_dropout_mask: the same hash), and all
    gradients equal those of the un-dropped op fed d out = d y * mask / keep_prob.  The stream is not ATen's Philox: parity GIVEN the mask."""
    from mmbidaf_b200 import functional as Fn, ops
    bsz, max_len, fan_in, hid, pr = cfg
    gen = torch.Generator().manual_seed(77 * bsz + max_len)
    lengths = torch.randint(1, max_len + 1, (bsz,), generator=gen).tolist()
    lengths[0] = max_len
    dev = "cuda"
    ws = [((torch.rand(*shape, generator=gen) - 0.5) * 0.2).to(dev) for _ in range(2)
          for shape in ((4 * hid, fan_in), (4 * hid, hid), (4 * hid,), (4 * hid,))]
    x = torch.randn(bsz, max_len, fan_in, generator=gen).to(dev)
    g_out = torch.randn(bsz, max_len, 2 * hid, generator=gen).to(dev)
    g_h = torch.randn(bsz, 2, hid, generator=gen).to(dev)
    len_d = torch.tensor(lengths, dtype=torch.int32, device=dev)
    ord_d = O.sort_order(lengths).to(torch.int32).to(dev)
    ops.rng_seed(1234)
    key = ops.rng_next_keys(dev, 3)[1:2]
    mask = ops.dropout_mask(key, 1.0 - pr, (bsz, max_len, 2 * hid))
    frac = mask.float().mean().item()
    assert abs(frac - (1.0 - pr)) < 4.0 * (pr * (1 - pr) / mask.numel()) ** 0.5 + 1e-3, frac      # the hash is a fair coin of the right bias

    def run(dropped):
        w = [t.clone().requires_grad_(True) for t in ws]
        xin = x.clone().requires_grad_(True)
        if dropped:
            y, h_n = Fn.lstm_layer(xin, len_d, ord_d, w, key, pr)
            ((y * g_out).sum() + (h_n * g_h).sum()).backward()
        else:
            out, h_n = Fn.lstm_layer(xin, len_d, ord_d, w)
            y = out * mask / (1.0 - pr)
            ((y * g_out).sum() + (h_n * g_h).sum()).backward()
        return y.detach(), h_n.detach(), xin.grad, [t.grad for t in w]

    y1, h1, dx1, dw1 = run(True)
    y0, h0, dx0, dw0 = run(False)
    assert torch.equal(h1, h0)
    assert rel_err(y1, y0) < 1e-6 and torch.equal(y1 == 0, y0 == 0)
    for b, n in enumerate(lengths):
        assert (y1[b, n:] == 0).all()
    assert grad_err(dx1, dx0) < 1e-5
    for a, b in zip(dw1, dw0):
        assert grad_err(a, b) < 1e-5


def test_in_kernel_dropout_keys_advance_and_encoder_uses_them():
    """Successive draws give different masks (also inside a replayed CUDA graph: the key state lives on the device); RNNEncoder in
    training mode drops both the inter-layer and the output activations, in eval mode nothing."""
    from mmbidaf_b200 import ops
    from mmbidaf_b200.layers import RNNEncoder
    dev = "cuda"
    ops.rng_seed(5)
    k = ops.rng_next_keys(dev, 2)
    k2 = ops.rng_next_keys(dev, 2)
    assert len({int(v) for v in torch.cat([k, k2]).tolist()}) == 4
    torch.manual_seed(0)
    enc = RNNEncoder(12, 16, 2, drop_prob=0.3).to(dev)
    x = torch.randn(6, 20, 12, device=dev)
    lengths = [20, 3, 17, 9, 20, 1]
    enc.eval()
    ref, ref_h = enc(x, lengths)
    enc.train()
    y1, h1 = enc(x, lengths)
    y2, _ = enc(x, lengths)
    valid = torch.zeros(6, 20, dtype=torch.bool, device=dev)
    for b, n in enumerate(lengths):
        valid[b, :n] = True
    zero1 = (y1 == 0)[valid].float().mean().item()
    assert 0.2 < zero1 < 0.4, zero1                                # ~30 % of the valid outputs dropped
    assert not torch.equal(y1 == 0, y2 == 0)                       # a fresh key per call
    assert not torch.allclose(h1, ref_h)                           # inter-layer dropout reaches layer 2's state
    assert (y1[~valid] == 0).all() and (ref[~valid] == 0).all()
    # graph replay: new masks every replay
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        enc(x, lengths)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        yg, _ = enc(x, lengths)
    g.replay()
    a = (yg == 0).clone()
    g.replay()
    b = (yg == 0).clone()
    assert not torch.equal(a, b)


@pytest.mark.parametrize("shape", [(3, 7, 10), (32, 409, 300), (5, 3), (1, 1, 1)])
def test_dropout_apply_and_mask_bytes(shape):
    """encoding.py:26 `F.dropout(x, drop_prob, training)` as one own launch: y = x * mask / keep_prob bit for bit with the mask of the
    same key; the backward applies the same mask to the gradient; BiDAFAttention's byte masks are the same bits."""
    from mmbidaf_b200 import functional as Fn, ops
    dev = "cuda"
    torch.manual_seed(3)
    x = torch.randn(*shape, device=dev, requires_grad=True)
    key = ops.rng_next_keys(dev, 1)
    y = Fn._Dropout.apply(x, key, 0.8)
    mask = ops.dropout_mask(key, 0.8, shape)
    assert torch.equal(y, torch.where(mask, x.detach() * (1.0 / 0.8), torch.zeros_like(x)))
    g = torch.randn_like(y)
    y.backward(g)
    assert torch.equal(x.grad, torch.where(mask, g * (1.0 / 0.8), torch.zeros_like(g)))
    assert torch.equal(ops.dropout_mask_u8(key, 0.8, shape).bool(), mask)
    assert Fn.dropout(x, 0.3, False) is x


def test_bidaf_layer_draws_own_masks():
    from mmbidaf_b200.layers import BiDAFAttention
    dev = "cuda"
    torch.manual_seed(1)
    att = BiDAFAttention(16, drop_prob=0.25).to(dev).train()
    kc, kq, scale = att._dropout_masks(torch.empty(4, 50, 16, device=dev), torch.empty(4, 30, 16, device=dev))
    assert kc.dtype == torch.uint8 and kc.shape == (4, 50, 16) and kq.shape == (4, 30, 16) and abs(scale - 1 / 0.75) < 1e-6
    assert 0.65 < kc.float().mean().item() < 0.85 and 0.65 < kq.float().mean().item() < 0.85
    assert not torch.equal(kc[:, :30], kq)                       # two keys
    att.eval()
    assert att._dropout_masks(torch.empty(1, 2, 16, device=dev), torch.empty(1, 2, 16, device=dev)) == (None, None, 1.0)
