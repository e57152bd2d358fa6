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
        with pytest.raises(ValueError, match="another bucket"):
            graphed.step_graphed(make_batch(5, 34, 70, 9, 4, seed=12).to("cuda"))
    finally:
        mmbidaf_b200.set_precision("fp32")


@pytest.mark.parametrize("tier", ["fp32", "fast"])
def test_graph_captured_on_one_batch_replays_other_batches_bit_identically(tier):
    """The collator pads every batch to its own maxima and lengths change every batch (datasets.py:298-302, train.py:126-146).
    A graph captured on batch A must serve any batch of the same padded-shape bucket: lengths, orders and masks are static
    device tensors rewritten in place (layers/encoding.py::LengthPlan).  Sequence A, B, C, B through the graph == eager."""
    import mmbidaf_b200
    from mmbidaf_b200.trainer import Trainer
    mmbidaf_b200.set_precision(tier)
    try:
        batches = [make_batch(5, 33, 70, 9, 4, seed=s).to("cuda") for s in (11, 12, 13)]
        assert batches[0].text_len != batches[1].text_len and batches[1].audio_len != batches[2].audio_len
        for b in batches[1:]:                           # same bucket: make_batch forces sample 0 to the maximum lengths
            assert b.text.shape == batches[0].text.shape and b.max_dec_len == batches[0].max_dec_len
        order = [0, 1, 2, 1]
        work = torch.cuda.Stream()
        with torch.cuda.stream(work):
            eager, graphed = Trainer(_make()), Trainer(_make())
            graphed.capture(batches[0], warmup=2)        # (two warm-up steps on batch A)
            for _ in range(2):
                eager.step(batches[0])
            losses_e = [eager.step(batches[i]).clone() for i in order]
            losses_g = [graphed.step_graphed(batches[i]).clone() for i in order]
            torch.cuda.synchronize()
        for le, lg in zip(losses_e, losses_g):
            assert torch.isfinite(le) and torch.equal(le, lg)
        assert len({float(v) for v in losses_e}) > 1       # the batches really differ
        for p, q in zip(eager.model.parameters(), graphed.model.parameters()):
            assert torch.equal(p, q)
    finally:
        mmbidaf_b200.set_precision("fp32")


def test_length_plan_kernel_bit_exact_masks_and_stable_order():
    """csrc/length_plan.cu against models.py:86-92 (get_mask) and :119-123 (decoder mask), and the stable longest-first order."""
    from mmbidaf_b200 import ops
    gen = torch.Generator().manual_seed(5)
    for B, L, M in ((1, 1, 0), (3, 7, 11), (32, 409, 409), (5, 33, 64), (64, 1024, 0), (257, 3, 5)):
        lens = torch.randint(0, L + 1, (B,), generator=gen)
        lens[0] = L
        dev = lens.to(torch.int32).cuda()
        mask, dec, order = ops.length_plan(dev, L, M, want_order=True)
        idx = torch.arange(L).unsqueeze(0).expand(B, L)
        want = idx < lens.unsqueeze(1)                                               # models.py:88-91
        assert mask.dtype == torch.bool and torch.equal(mask.cpu(), want)
        if M:
            want_dec = torch.cat((want, torch.zeros(B, M - L).type(want.type())), dim=1)   # models.py:121-123
            assert torch.equal(dec.cpu(), want_dec)
        else:
            assert dec is None
        stable = sorted(range(B), key=lambda i: -int(lens[i]))
        assert order.cpu().tolist() == stable


def test_col_sum_any_width():
    from mmbidaf_b200 import ops
    gen = torch.Generator().manual_seed(9)
    for n, p in ((1000, 10), (33, 7), (4096, 800), (5, 1), (70000, 12)):
        a = torch.randn(n, p, generator=gen).cuda()
        got = ops.col_sum(a)
        want = a.double().sum(dim=0)
        assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-4


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


@pytest.mark.parametrize("wd", [0.0, 0.01])
@pytest.mark.parametrize("gscale", [1e-3, 5.0])        # norm below / above max_grad_norm: un-clipped and clipped
def test_fused_clip_adadelta_matches_torch_semantics(wd, gscale):
    """csrc/optimizer.cu against clip_grad_norm_ + torch.optim.Adadelta (train.py:154-155, :110) in fp64, three steps."""
    from mmbidaf_b200 import ops
    n, lr, rho, eps, max_norm = 40_004, 0.5, 0.9, 1e-6, 2.0
    gen = torch.Generator().manual_seed(7)
    p0 = torch.randn(n, generator=gen)
    ref = torch.nn.Parameter(p0.double().clone())
    opt = torch.optim.Adadelta([ref], lr=lr, rho=rho, eps=eps, weight_decay=wd)
    param = p0.cuda()
    sq, acc = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(3):
        g = torch.randn(n, generator=gen) * gscale
        ref.grad = g.double().clone()
        torch.nn.utils.clip_grad_norm_([ref], max_norm)
        opt.step()
        grad = g.cuda()
        norm = torch.linalg.vector_norm(grad)
        ops.adadelta_clip_step(param, grad, sq, acc, norm, max_norm, lr, rho, eps, wd)
        torch.cuda.synchronize()
        assert torch.allclose(grad.cpu().double(), ref.grad, rtol=1e-5, atol=1e-9)            # clipped gradient written back
        assert torch.allclose(param.cpu().double(), ref.detach(), rtol=1e-5, atol=1e-6)
    st = opt.state[ref]
    assert torch.allclose(sq.cpu().double(), st["square_avg"], rtol=1e-5, atol=1e-12)
    assert torch.allclose(acc.cpu().double(), st["acc_delta"], rtol=1e-4, atol=1e-12)


def test_second_backward_over_one_forward_raises():
    """The LSTM backward overwrites its saved gates in place and the decoder tape is consumed by its closing backward: a second
    backward over the same graph must raise, not return silently wrong gradients."""
    import pytest as _pytest
    from mmbidaf_b200.layers import RNNEncoder
    torch.manual_seed(0)
    enc = RNNEncoder(8, 6, 1).cuda()
    x = torch.randn(3, 5, 8, device="cuda", requires_grad=True)
    out, _ = enc(x, [5, 3, 2])
    loss = out.sum()
    loss.backward(retain_graph=True)
    with _pytest.raises(RuntimeError, match="ran twice"):
        loss.backward()
