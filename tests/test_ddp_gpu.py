"""GPU, 2 ranks over NCCL: the real MMBiDAF, data parallel (one process per GPU, one flat all-reduce(SUM) of the gradient,
trainer.py) reproduces the single-process update of the global batch -- with ragged text lengths, so that a shard's own
maximum text length is below the global one (the case shard_batch must not re-pad).  Skipped with < 2 GPUs."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmbidaf_b200.synth import make_batch

pytestmark = pytest.mark.gpu
DIMS = (100, 300, 128, 1000, 64)


def _model(seed=3):
    from mmbidaf_b200.models import MMBiDAF
    torch.manual_seed(seed)
    return MMBiDAF(*DIMS[:4], torch.device("cuda"), drop_prob=0.0, max_transcript_length=DIMS[4]).cuda()


def _batch():
    b = make_batch(6, 40, 90, 11, 5, seed=31)
    # the global maxima live in rank 0's shard only: rank 1's own text / target maxima are smaller
    for i in (3, 4, 5):
        n = min(b.text_len[i], 23)
        b.text[i, n - 1:] = 0
        b.text[i, n - 1] = -1.0
        b.text_len[i] = n
        k = min(b.target_len[i], 3)
        b.targets[i, k - 1:] = 0
        b.targets[i, k - 1, 0] = float(n - 1)
        b.targets[i, :k - 1, 0] = b.targets[i, :k - 1, 0].clamp(max=n - 2)
        b.target_len[i] = k
    return b


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mmbidaf_b200.trainer import Trainer, shard_batch
    trainer = Trainer(_model(seed=3 + rank))             # replicas differ until Trainer broadcasts rank 0's weights
    shard = shard_batch(_batch(), rank, world).to("cuda")
    losses = [float(trainer.step(shard)) for _ in range(2)]
    torch.cuda.synchronize()
    out[rank] = ([p.detach().cpu() for p in trainer.model.parameters()], float(trainer.last_grad_norm), losses)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_nccl_ranks_equal_the_single_process_global_batch():
    from mmbidaf_b200.trainer import Trainer
    world, port = 2, 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        results = {r: out[r] for r in range(world)}
    single = Trainer(_model(seed=3))
    batch = _batch().to("cuda")
    losses = [float(single.step(batch)) for _ in range(2)]
    want = [p.detach().cpu() for p in single.model.parameters()]
    for step in range(2):                                # the loss is a sum over videos (models.py:170): rank losses add up
        assert abs(results[0][2][step] + results[1][2][step] - losses[step]) <= 1e-4 * abs(losses[step])
    assert abs(results[0][1] - float(single.last_grad_norm)) <= 1e-4 * float(single.last_grad_norm)
    for a, b in zip(results[0][0], results[1][0]):
        assert torch.equal(a, b)                         # replicas stay bit-identical
    worst = max(float((g - w).abs().max()) / max(float(w.abs().max()), 1e-6) for g, w in zip(results[0][0], want))
    assert worst < 1e-4, worst


def test_sharded_gradients_sum_to_the_global_gradient_on_one_gpu():
    """The same equivalence without NCCL (runs on the 1-GPU box): gradients of the two shards, summed, against the gradient of
    the global batch -- fails if shard_batch re-pads the text (un-masked decoder soft-maxes, attention.py:148,154)."""
    from mmbidaf_b200.trainer import shard_batch
    model = _model()
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]

    def grads(b):
        b = b.to("cuda")
        _, loss = model(b.text, b.text_len, b.audio, b.audio_len, b.images, b.image_len, b.targets, b.target_len, b.max_dec_len)
        return loss.detach(), torch.autograd.grad(loss, params, allow_unused=True)

    loss, want = grads(_batch())
    parts = [grads(shard_batch(_batch(), r, 2)) for r in range(2)]
    assert abs(float(parts[0][0] + parts[1][0] - loss)) <= 1e-5 * abs(float(loss))
    for w, g0, g1 in zip(want, parts[0][1], parts[1][1]):
        if w is None:
            continue
        err = float((g0 + g1 - w).abs().max()) / max(float(w.abs().max()), 1e-3)
        assert err < 2e-4, err
