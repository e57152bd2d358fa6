"""CPU: the data-parallel training step logic (sharding, flat-bucket SUM all-reduce, global-norm clip,
Adadelta) under gloo with world_size 2, on a small stand-in model with the MMBiDAF call signature."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmbidaf_b200.synth import make_batch
from mmbidaf_b200.trainer import FlatGrads, Trainer, shard_batch, shard_range


class TinyModel(torch.nn.Module):
    """Same forward signature / loss convention as MMBiDAF (loss summed over the batch)."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.t = torch.nn.Linear(300, 4)
        self.a = torch.nn.Linear(128, 4)

    def forward(self, text, text_len, audio, audio_len, images, image_len, targets, target_len, max_dec_len):
        loss = text.new_zeros(())
        for b, (n, m) in enumerate(zip(text_len, audio_len)):          # per-video terms, padding never visited
            loss = loss + self.t(text[b, :n]).tanh().sum() + self.a(audio[b, :m]).sigmoid().sum()
        return None, loss


def test_shard_range_is_a_partition():
    for n in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            seen = [i for r in range(world) for i in shard_range(n, r, world)]
            assert seen == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_keeps_text_and_target_padding_of_the_global_batch():
    """Audio / key-frames are re-padded to the shard's own maxima (masked everywhere); text and targets keep the global widths:
    the decoder's un-masked attention soft-maxes (attention.py:148,154) and the padded decode steps (models.py:168,179) see them."""
    batch = make_batch(5, 12, 20, 4, 3, seed=9)
    parts = [shard_batch(batch, r, 2) for r in range(2)]
    assert sum(len(p.text_len) for p in parts) == 5
    for p in parts:
        assert p.text.shape[1] == batch.text.shape[1] and p.targets.shape[1] == batch.targets.shape[1]
        assert p.max_dec_len == batch.max_dec_len
        assert p.audio.shape[1] == max(p.audio_len) and p.images.shape[1] == max(p.image_len)
    assert torch.equal(parts[1].text[0], batch.text[3]) and torch.equal(parts[1].targets[0], batch.targets[3])


def test_flat_grads_clip_matches_torch():
    model = TinyModel()
    ref = TinyModel()
    fg = FlatGrads(model.parameters())
    batch = make_batch(3, 6, 7, 2, 2, seed=2)
    for m in (model, ref):
        m(batch.text, batch.text_len, batch.audio, batch.audio_len, None, None, None, None, None)[1].backward()
    want = torch.nn.utils.clip_grad_norm_(ref.parameters(), 2.0)
    got = fg.clip_(2.0)
    assert torch.allclose(got, want)
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-8)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = TinyModel()
    if rank == 1:                                       # replicas must be re-synchronised from rank 0
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    trainer = Trainer(model)
    batch = make_batch(5, 12, 20, 4, 3, seed=9)
    for _ in range(2):
        trainer.step(shard_batch(batch, rank, world))
    out[rank] = [p.detach().clone() for p in model.parameters()] + [trainer.last_grad_norm.clone()]
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_global_batch():
    world, port = 2, 29000 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        results = {r: out[r] for r in range(world)}
    single = TinyModel()
    trainer = Trainer(single)
    batch = make_batch(5, 12, 20, 4, 3, seed=9)
    for _ in range(2):
        trainer.step(batch)
    want = [p.detach() for p in single.parameters()] + [trainer.last_grad_norm]
    for r in range(world):
        for got, w in zip(results[r], want):
            assert torch.allclose(got, w, rtol=1e-5, atol=1e-7)
    for a, b in zip(results[0], results[1]):
        assert torch.equal(a, b)                         # replicas bit-identical after the update


def test_flat_adadelta_matches_torch_optimizer():
    from mmbidaf_b200.trainer import FlatAdadelta, FlatState
    a, b = TinyModel(), TinyModel()
    batch = make_batch(3, 6, 7, 2, 2, seed=4)
    state = FlatState(a.parameters())
    mine = FlatAdadelta(state, lr=0.5, weight_decay=0.01)
    ref = torch.optim.Adadelta(b.parameters(), lr=0.5, weight_decay=0.01)
    for _ in range(3):
        state.zero()
        ref.zero_grad()
        for m in (a, b):
            m(batch.text, batch.text_len, batch.audio, batch.audio_len, None, None, None, None, None)[1].backward()
        mine.step()
        ref.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-7)


def test_flat_state_puts_lstm_directions_side_by_side():
    """trainer.FlatState orders the two directions of every nn.LSTM tensor next to each other, so that the recurrence's stacked
    operands (functional._stacked) are views of the flat parameter buffer -- and stay copies for parameters that are not adjacent."""
    import torch.nn as nn
    from mmbidaf_b200.functional import _stacked
    from mmbidaf_b200.trainer import FlatState

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.head = nn.Linear(12, 7)
            self.rnn = nn.LSTM(12, 8, 2, batch_first=True, bidirectional=True)
            self.tail = nn.Linear(3, 5)

    torch.manual_seed(0)
    net = Net()
    before = {k: v.detach().clone() for k, v in net.named_parameters()}
    state = FlatState(net.parameters())
    for k, v in net.named_parameters():
        assert torch.equal(v, before[k])                         # values kept, every parameter a view of the flat buffer
        assert v.untyped_storage().data_ptr() == state.flat_param.untyped_storage().data_ptr()
    for layer, fan_in in ((0, 12), (1, 16)):
        f, r = getattr(net.rnn, f"weight_ih_l{layer}"), getattr(net.rnn, f"weight_ih_l{layer}_reverse")
        w = _stacked([f, r], (2 * 32, fan_in))
        assert w.data_ptr() == f.data_ptr() and torch.equal(w, torch.cat([f, r]))          # a view, same values
        hf, hr = getattr(net.rnn, f"weight_hh_l{layer}"), getattr(net.rnn, f"weight_hh_l{layer}_reverse")
        wh = _stacked([hf, hr], (2, 32, 8))
        assert wh.data_ptr() == hf.data_ptr() and torch.equal(wh, torch.stack([hf, hr]))
        bf, br = getattr(net.rnn, f"bias_hh_l{layer}"), getattr(net.rnn, f"bias_hh_l{layer}_reverse")
        assert _stacked([bf, br], (64,)).data_ptr() == bf.data_ptr()
    # a two-layer highway encoder: gate and transform of a layer side by side (weights, then biases)
    from mmbidaf_b200.layers import HighwayEncoder
    hwy = HighwayEncoder(2, 8)
    FlatState(hwy.parameters())
    for k in range(2):
        gw, tw, gb, tb = hwy.gates[k].weight, hwy.transforms[k].weight, hwy.gates[k].bias, hwy.transforms[k].bias
        w = _stacked([gw, tw], (16, 8))
        assert w.data_ptr() == gw.data_ptr() and torch.equal(w, torch.cat([gw, tw]))
        assert _stacked([gb, tb], (16,)).data_ptr() == gb.data_ptr()
    a, b = torch.randn(4, 3), torch.randn(4, 3)                    # unrelated tensors: a copy
    c = _stacked([a, b], (8, 3))
    assert c.data_ptr() != a.data_ptr() and torch.equal(c, torch.cat([a, b]))
