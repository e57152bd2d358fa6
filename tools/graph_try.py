"""Does the whole training step capture into a CUDA graph, and what does replay cost?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, mmbidaf_b200
from mmbidaf_b200.models import MMBiDAF
from mmbidaf_b200.synth import make_batch
from mmbidaf_b200.trainer import Trainer
mmbidaf_b200.set_precision("fast")
dev = torch.device("cuda:0")
torch.manual_seed(224)
model = MMBiDAF(bench.HIDDEN, bench.E_TEXT, bench.E_AUDIO, bench.E_IMAGE, dev, drop_prob=0.0, max_transcript_length=bench.M).to(dev)
use_streams = "--no-streams" not in sys.argv
model.use_streams = use_streams
tr = Trainer(model)
c = bench.CFG3
batch = make_batch(c["batch"], c["lt"], c["la"], c["li"], c["t_dec"], seed=224).to(dev)
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    l_eager = tr.step(batch)
torch.cuda.synchronize()
print("eager ms/step", (time.perf_counter() - t0) * 100, float(l_eager))
torch.manual_seed(224)
model2 = MMBiDAF(bench.HIDDEN, bench.E_TEXT, bench.E_AUDIO, bench.E_IMAGE, dev, drop_prob=0.0, max_transcript_length=bench.M).to(dev)
model2.use_streams = use_streams
tr2 = Trainer(model2)
for _ in range(13 + 3 + 10):
    l_ref = tr2.step(batch)          # eager twin: same number of updates as eager(13) + capture warm-up(3) + 10 replays
mode = [a.split("=")[1] for a in sys.argv if a.startswith("--mode=")]
tr.capture(batch, error_mode=mode[0] if mode else "global")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    l_graph = tr.step_graphed()
torch.cuda.synchronize()
print("graph ms/step", (time.perf_counter() - t0) * 100, float(l_graph), "streams", use_streams)
diff = max(float((p - q).abs().max()) for p, q in zip(model.parameters(), model2.parameters()))
print("twin loss", float(l_ref), "max param diff vs eager twin", diff)
