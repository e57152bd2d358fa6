"""Stage timeline of the one-kernel decoder BACKWARD step (csrc/decoder_fused.cu: dec_bwd_head_kernel with the sweeps), block 0:
clock64 stamps at the stage boundaries of the last backward step that ran (the first forward step of the sequence).
    python tools/decoder_bwd_trace.py [B Lt]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B, Lt = (int(v) for v in sys.argv[1:3]) if len(sys.argv) >= 3 else (32, 409)
H, E, M, STEPS = 100, 300, 409, 4
dev = "cuda"
tr = torch.zeros(32, dtype=torch.int64, device=dev)
os.environ["MMB_DEC_TRACE"] = str(tr.data_ptr())
from mmbidaf_b200.layers import MultimodalAttentionDecoder  # noqa: E402

torch.manual_seed(0)
dec = MultimodalAttentionDecoder(E, H, M).to(dev).train()
enc_a = torch.randn(B, Lt, 2 * H, device=dev, requires_grad=True)
enc_i = torch.randn(B, Lt, 2 * H, device=dev, requires_grad=True)
mask = torch.ones(B, M, dtype=torch.bool, device=dev)
h = torch.randn(B, 1, H, device=dev, requires_grad=True)
cell, cov = torch.zeros(1, B, H, device=dev), torch.zeros(B, Lt, 1, device=dev)
loss = 0
for k in range(STEPS):
    sent = torch.randn(B, 1, E, device=dev)
    tgt = torch.randint(0, M, (B,), device=dev)
    probs, h, cell, att, cov, terms = dec.step(sent, h, cell, enc_a, enc_i, cov, mask, target=tgt)
    loss = loss + terms.sum()
loss.backward()
torch.cuda.synchronize()
t = tr.cpu().tolist()
names = ["1 masked soft-max bwd", "2 d h = d_logits out.weight", "3 LSTM cell bwd", "4 d c3 mat-vec", "5 modality soft-max / W_beta bwd",
         "6 d c_k mat-vec", "7 sweep: d alpha (enc rows)", "8 sweep: soft-max / tanh bwd, d proj", "9 column sums, d h mat-vec"]
for i, nm in enumerate(names):
    print(f"{nm:40s} {t[19 + i] - t[18 + i]:7d} cycles")
print(f"{'total':40s} {t[27] - t[18]:7d} cycles = {(t[27] - t[18]) / 1.965e3:.1f} us")
