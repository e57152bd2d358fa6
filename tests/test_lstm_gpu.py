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
