"""Which part of the step refuses CUDA-graph capture?"""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmbidaf_b200
from mmbidaf_b200.layers import BiDAFAttention, RNNEncoder, MultimodalAttentionDecoder, Embedding
dev = "cuda"
if os.environ.get('PREC'): mmbidaf_b200.set_precision(os.environ['PREC'])
BIG = os.environ.get('BIG') == '1'

def try_capture(name, fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name, flush=True)
    except Exception as e:
        print("FAIL", name, str(e).splitlines()[0][:120], flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass

B, L, H = 4, 20, 100
lens = [20, 13, 9, 20]
x = torch.randn(B, L, H, device=dev)
enc = RNNEncoder(H, H, 1).to(dev).train()
def lstm_fwd():
    with torch.no_grad(): enc(x, lens)
def lstm_fb():
    out, h = enc(x, lens)
    torch.autograd.grad(out.sum() + h.sum(), list(enc.parameters()))
emb = Embedding(300, H, 0.0).to(dev)
xe = torch.randn(B, L, 300, device=dev)
def emb_fb():
    torch.autograd.grad(emb(xe).sum(), list(emb.parameters()))
bid = BiDAFAttention(2 * H, 0.0).to(dev).train()
c = torch.randn(B, L, 2 * H, device=dev, requires_grad=True); q = torch.randn(B, 12, 2 * H, device=dev, requires_grad=True)
cm = torch.ones(B, L, dtype=torch.bool, device=dev); qm = torch.ones(B, 12, dtype=torch.bool, device=dev)
def bidaf_fb():
    torch.autograd.grad(bid(c, q, cm, qm).sum(), [c, q] + list(bid.parameters()))
dec = MultimodalAttentionDecoder(300, H, 30).to(dev).train()
mask = torch.ones(B, 30, dtype=torch.bool, device=dev)
def dec_fb():
    ea, ei = c * 1.0, c * 2.0
    h = torch.zeros(B, 1, H, device=dev); cell = torch.zeros(1, B, H, device=dev); cov = torch.zeros(B, L, 1, device=dev)
    sent = torch.zeros(B, 1, 300, device=dev); loss = 0
    for s in range(2):
        p, h, cell, att, cov = dec(sent, h, cell, ea, ei, cov, mask)
        loss = loss - torch.log(p[:, s] + 1e-12).sum() + torch.min(att, cov).sum()
    torch.autograd.grad(loss, [c] + list(dec.parameters()), allow_unused=True)
import bench
from mmbidaf_b200.models import MMBiDAF
from mmbidaf_b200.synth import make_batch
from mmbidaf_b200.trainer import Trainer
model = MMBiDAF(100, 300, 128, 1000, torch.device(dev), drop_prob=0.0, max_transcript_length=409 if BIG else 40).to(dev).train()
model.use_streams = False
bt = (make_batch(32, 409, 1024, 128, 12, seed=1) if BIG else make_batch(4, 20, 30, 6, 3, seed=1)).to(dev)
def model_fwd():
    with torch.no_grad():
        model(bt.text, bt.text_len, bt.audio, bt.audio_len, bt.images, bt.image_len, bt.targets, bt.target_len, bt.max_dec_len)
def model_fb():
    _, loss = model(bt.text, bt.text_len, bt.audio, bt.audio_len, bt.images, bt.image_len, bt.targets, bt.target_len, bt.max_dec_len)
    torch.autograd.grad(loss, [p for p in model.parameters() if p.requires_grad], allow_unused=True)
tr = Trainer(model)
def full_step():
    tr.step(bt)
def model_fb_streams():
    model.use_streams = True
    model_fb()
for name, fn in [("model fwd", model_fwd), ("model fwd+bwd", model_fb), ("trainer step", full_step), ("model fwd+bwd streams", model_fb_streams),("lstm fwd", lstm_fwd), ("emb fwd+bwd", emb_fb), ("lstm fwd+bwd", lstm_fb), ("bidaf fwd+bwd", bidaf_fb), ("decoder fwd+bwd", dec_fb)]:
    try_capture(name, fn)
