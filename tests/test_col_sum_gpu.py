"""GPU: column sums of tall matrices (the bias gradients of the training step) against an fp64 sum."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 4), (3, 8), (63, 8), (64, 128), (65, 132), (1000, 200), (13088, 800), (32768, 800),
                                   (32768, 200), (4097, 1000), (7, 409)])
def test_col_sum_matches_fp64(shape):
    from mmbidaf_b200 import ops
    n, p = shape
    gen = torch.Generator().manual_seed(n * 31 + p)
    a = torch.randn(n, p, generator=gen) + 0.25
    got = ops.col_sum(a.cuda())
    torch.cuda.synchronize()
    want = a.double().sum(dim=0)
    scale = a.double().abs().sum(dim=0)
    assert got.shape == (p,)
    assert ((got.cpu().double() - want).abs() <= 2e-6 * scale).all()     # fp32 tree sum: ~1e-7 relative to sum |a|
    again = ops.col_sum(a.cuda())                                        # deterministic: no atomics
    assert torch.equal(got, again)


def test_col_sum_blocks_is_host_only_and_bounded():
    from mmbidaf_b200 import _lib
    lib = _lib.lib()
    assert lib.mmb_col_sum_blocks(1, 4) == 1
    assert lib.mmb_col_sum_blocks(32768, 800) * 7 <= 4 * 148 + 7
    assert lib.mmb_col_sum_blocks(0, 4) == 0
