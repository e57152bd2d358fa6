"""CPU: host-side logic that needs no GPU (greedy decode, synthetic batches, parameter naming)."""
import torch

from mmbidaf_b200.decode import greedy_search
from mmbidaf_b200.synth import make_batch
from oracle import mmbidaf_oracle as O


def test_greedy_search_matches_oracle():
    gen = torch.Generator().manual_seed(3)
    for text_len in (2, 5, 9):
        dist = torch.rand(7, 12, generator=gen)
        assert greedy_search(dist, text_len) == O.greedy_indices(dist, text_len)


def test_synthetic_batch_contract():
    b = make_batch(4, 20, 30, 6, 5, seed=1)
    assert b.text.shape == (4, 20, 300) and b.audio.shape == (4, 30, 128) and b.images.shape == (4, 6, 1000, 1, 1)
    assert b.text_len[0] == 20 and max(b.text_len) == 20 and min(b.text_len) >= 10
    for i, n in enumerate(b.text_len):
        assert (b.text[i, n - 1] == -1).all() and (b.text[i, n:] == 0).all()        # EOS row, zero padding
        k = b.target_len[i]
        assert b.targets[i, k - 1, 0] == n - 1 and (b.targets[i, k:] == 0).all()
        assert (b.targets[i, :k - 1, 0] < n - 1).all()
    assert b.max_dec_len == max(b.target_len)


def test_model_parameter_names_match_reference_layout():
    from mmbidaf_b200.models import MMBiDAF
    m = MMBiDAF(6, 10, 4, 12, torch.device("cpu"), max_transcript_length=11)
    want = O.param_shapes(6, 10, 4, 12, 11)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    n_params = sum(p.numel() for p in MMBiDAF(100, 300, 128, 1000, torch.device("cpu"), max_transcript_length=409).parameters())
    assert n_params == 3201715          # trainable parameters of the reference model


def test_layers_refuse_cpu_tensors():
    import pytest
    from mmbidaf_b200.layers import BiDAFAttention, RNNEncoder
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BiDAFAttention(8)(torch.zeros(1, 2, 8), torch.zeros(1, 2, 8), torch.ones(1, 2, dtype=torch.bool),
                          torch.ones(1, 2, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RNNEncoder(4, 4, 1)(torch.zeros(1, 2, 4), [2])


def test_decode_loss_matches_the_reference_formula_and_its_gradient():
    """functional.decode_loss against models.py:168-179 (training: coverage term of every step) and :197-199 (evaluation:
    of the last step only), values and the gradient handed to every step (host tensors: the function is pure torch)."""
    import pytest
    from mmbidaf_b200 import functional as Fn
    gen = torch.Generator().manual_seed(3)
    steps, B, w = 5, 4, 1.0
    for every_step in (True, False):
        terms = [torch.rand(2, B, generator=gen).double().requires_grad_(True) for _ in range(steps)]
        ref_terms = [t.detach().clone().requires_grad_(True) for t in terms]
        got = Fn.decode_loss(terms, steps, w, every_step)
        stacked = torch.stack(ref_terms)                                  # (steps, 2, B)
        cov = stacked[:, 1].sum() if every_step else stacked[-1, 1].sum()
        want = (stacked[:, 0].sum() + w * cov) / steps
        assert float(got) == pytest.approx(float(want), rel=1e-12)
        (3.0 * got).backward()
        (3.0 * want).backward()
        for t, r in zip(terms, ref_terms):
            assert torch.allclose(t.grad, r.grad, rtol=1e-6, atol=0)      # the shared coefficient tensor is fp32


def test_leaf_lanes_is_a_no_op_without_cuda_work():
    from mmbidaf_b200 import functional as Fn
    with Fn.leaf_lanes():
        with Fn._leaf(torch.zeros(3)):
            x = torch.ones(2) + 1
    assert x.tolist() == [2.0, 2.0] and not Fn._Lanes.enabled


def test_tall_tn_split_reduction_matches_the_plain_product():
    """functional.tall_tn (every weight gradient of the step) splits the reduction over row blocks when the output has too
    few tiles; both paths against a^T b in fp64."""
    from mmbidaf_b200 import functional as Fn
    gen = torch.Generator().manual_seed(5)
    for n, p, q in ((4096, 400, 100), (1024, 200, 100), (1000, 64, 32), (130, 8, 8), (2048, 800, 800)):
        a, b = torch.randn(n, p, generator=gen), torch.randn(n, q, generator=gen)
        got = Fn.tall_tn(a, b)
        want = a.double().t() @ b.double()
        assert got.shape == (p, q)
        assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-4
