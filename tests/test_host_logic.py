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
