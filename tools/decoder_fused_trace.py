"""Stage timeline of the one-kernel decoder step (csrc/decoder_fused.cu), block 0: clock64 stamps at the stage boundaries.
    python tools/decoder_fused_trace.py [B Lt]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B, Lt = (int(v) for v in sys.argv[1:3]) if len(sys.argv) >= 3 else (32, 409)
H, E, M = 100, 300, 409
dev = "cuda"
tr = torch.zeros(32, dtype=torch.int64, device=dev)
os.environ["MMB_DEC_TRACE"] = str(tr.data_ptr())
from mmbidaf_b200.layers import MultimodalAttentionDecoder  # noqa: E402

torch.manual_seed(0)
dec = MultimodalAttentionDecoder(E, H, M).to(dev).eval()
enc_a, enc_i = torch.randn(B, Lt, 2 * H, device=dev), torch.randn(B, Lt, 2 * H, device=dev)
mask = torch.ones(B, M, dtype=torch.bool, device=dev)
sent, h = torch.randn(B, 1, E, device=dev), torch.randn(B, 1, H, device=dev)
cell, cov = torch.zeros(1, B, H, device=dev), torch.zeros(B, Lt, 1, device=dev)
with torch.no_grad():
    for _ in range(4):
        probs, h, cell, att, cov = dec(sent, h, cell, enc_a, enc_i, cov, mask)
torch.cuda.synchronize()
t = tr.cpu().tolist()
names = ["inputs", "A hw mat-vec", "sync", "B energies", "B soft-max", "B contexts", "sync", "merge", "C pb mat-vec", "sync",
         "C beta / coverage", "D gates + cell", "sync", "E logits + local soft-max", "sync", "E probs / arg-max", "sync"]
for i, nm in enumerate(names):
    print(f"{nm:28s} {t[i + 1] - t[i]:7d} cycles")
print(f"{'total':28s} {t[17] - t[0]:7d} cycles = {(t[17] - t[0]) / 1.965e3:.1f} us")
