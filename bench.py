#!/usr/bin/env python
"""bench.py -- MMBiDAF training throughput on B200 (BASELINE.json metric) + fused-BiDAF roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W]                # our arm (N>1: launched by torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]          # the reference algorithm on the host CPU

One "step" = one pass of the hot path over one batch of synthetic Coursera-shaped videos:
forward -> loss.backward() -> gradient all-reduce(SUM) -> clip_grad_norm_(2.0) -> Adadelta(lr 0.5).step()
(reference train.py:146-155) on BASELINE config 3 (B=32 videos per GPU, Lt<=409, La<=1024, Li<=128,
T_dec<=12, hidden 100, text 300 / audio 128 / image 1000, drop 0.2).  At N GPUs every rank runs the same
per-GPU batch (weak scaling; N=8 is BASELINE config 4's global batch of 256).

One JSON line on stdout (rank 0).  value = videos/s with inputs resident in HBM; e2e = the same step fed
from pinned host memory with the loss read back every step; roofline = the fused BiDAF forward on BASELINE
config 2 (B=64, Lc=512, Lq=256, d=200) timed with CUDA events in this run against the measured HBM peak;
cpu_baseline = the oracle port of the reference algorithm timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, E_TEXT, E_AUDIO, E_IMAGE, M = 100, 300, 128, 1000, 409
CFG3 = dict(batch=32, lt=409, la=1024, li=128, t_dec=12)            # per GPU
CFG2 = dict(batch=64, lc=512, lq=256, d=200)
DROP = 0.2                                                           # train.py:209
METRIC, UNIT = "mmbidaf_train_videos_per_s", "videos/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = torch.cuda.get_device_properties(index).uuid
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:                                           # pragma: no cover
            self.nv, self.error = None, repr(exc)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.reasons.update(k for k, v in names.items() if bits & v)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the reference is Python and cannot travel to the
# GPU box).  The only place besides tests/ and smoke() that executes oracle/.
# ------------------------------------------------------------------------------------------------------
def cpu_training_throughput(videos_per_step: int, steps: int, warmup: int):
    from mmbidaf_b200.synth import make_batch
    from oracle import mmbidaf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = {k: v.clone().requires_grad_(True) for k, v in O.make_params(HIDDEN, E_TEXT, E_AUDIO, E_IMAGE, M).items()}
    plist = list(params.values())
    opt = torch.optim.Adadelta(plist, lr=0.5)
    batch = make_batch(videos_per_step, CFG3["lt"], CFG3["la"], CFG3["li"], CFG3["t_dec"], E_TEXT, E_AUDIO, E_IMAGE, seed=224)
    img = batch.images.flatten(2)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        _, loss = O.mmbidaf_forward(params, batch.text, batch.text_len, batch.audio, batch.audio_len, img,
                                    batch.image_len, batch.targets, batch.max_dec_len, M, training=True, fast_lstm=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(plist, 2.0)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    sample = (f"{videos_per_step} videos/step of config 3 shapes (Lt<=409, La<=1024, Li<=128, T_dec<=12), "
              f"{steps} timed + {warmup} warm-up steps, fwd+bwd+clip+Adadelta, fp32, dropout off, torch CPU {torch.__version__}")
    return videos_per_step * steps / total, total / steps * 1e3, cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vps, ms, cores, sample = cpu_training_throughput(4, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": round(vps, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": round(vps, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(vps, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int):
    return {"workload": "BASELINE config 3: full MMBiDAF training step (embedding->RNNEncoder->BiDAFx2->mod-LSTM->"
                        "multimodal attention decoder->loss; backward; clip 2.0; Adadelta lr 0.5)",
            "per_gpu_batch": CFG3["batch"], "global_batch": CFG3["batch"] * n_gpus, "max_text_len": CFG3["lt"],
            "max_audio_len": CFG3["la"], "max_image_len": CFG3["li"], "max_dec_len": CFG3["t_dec"], "hidden": HIDDEN,
            "embed": [E_TEXT, E_AUDIO, E_IMAGE], "max_transcript_length": M, "drop_prob": DROP,
            "parallelism": f"dp{n_gpus}", "launch": "whole step captured once into a CUDA graph and replayed", "l2": "inputs+activations per step exceed L2 (126 MB); BiDAF microbench rotates 4 input sets"}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def bidaf_microbench(device, iters: int, warmup: int, precision: int, backward: bool = False):
    """Fused BiDAF forward (or backward) on BASELINE config 2; returns (avg seconds per call, algorithmic bytes)."""
    from mmbidaf_b200 import ops
    B, Lc, Lq, d = CFG2["batch"], CFG2["lc"], CFG2["lq"], CFG2["d"]
    gen = torch.Generator().manual_seed(224)
    sets = []
    for _ in range(4):                                      # 4 x 144 MB > L2: every timed forward reads cold data
        c = torch.randn(B, Lc, d, generator=gen).to(device)
        q = torch.randn(B, Lq, d, generator=gen).to(device)
        c_len = torch.randint(Lc // 2, Lc + 1, (B,), generator=gen)
        q_len = torch.randint(Lq // 2, Lq + 1, (B,), generator=gen)
        cm = (torch.arange(Lc).unsqueeze(0) < c_len.unsqueeze(1)).to(device)
        qm = (torch.arange(Lq).unsqueeze(0) < q_len.unsqueeze(1)).to(device)
        sets.append((c, q, cm, qm))
    w = [torch.randn(d, generator=gen).to(device) * 0.1 for _ in range(3)]
    bias = torch.zeros(1, device=device)
    if backward:                                             # saved forward state + an upstream gradient per input set
        full = []
        for s in sets:
            out, q2c, lse_r, lse_c, bm, ws = ops.bidaf_fwd(s[0], s[1], s[2], s[3], w[0], w[1], w[2], bias,
                                                            precision=precision, save=True)
            full.append(s + (torch.randn(out.shape, generator=gen).to(device), out, q2c, lse_r, lse_c, bm, ws))
        sets = full
        run = lambda s: ops.bidaf_bwd(s[4], s[0], s[1], s[2], s[3], w[0], w[1], w[2], bias, None, None, 1.0, s[5], s[9], s[6], s[7],
                                      s[8], s[10], precision)
    else:
        run = lambda s: ops.bidaf_fwd(s[0], s[1], s[2], s[3], w[0], w[1], w[2], bias, precision=precision)
    for i in range(warmup):
        run(sets[i % 4])
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for i in range(iters):
        run(sets[i % 4])
    end.record()
    torch.cuda.synchronize()
    algo_bytes = 4 * B * (Lc * d + Lq * d + Lc * 4 * d) + B * (Lc + Lq)
    if backward:                                            # read dX (4 Lc d), c, q; write dc, dq
        algo_bytes = 4 * B * (4 * Lc * d + 2 * Lc * d + 2 * Lq * d)
    return start.elapsed_time(end) / 1e3 / iters, algo_bytes


def run_gpu_arm(args):
    import torch.distributed as dist
    from mmbidaf_b200 import ops
    from mmbidaf_b200.models import MMBiDAF
    from mmbidaf_b200.synth import make_batch
    from mmbidaf_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout (NCCL_DEBUG=VERSION in this image): keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    import mmbidaf_b200
    mmbidaf_b200.set_precision(args.precision)

    torch.manual_seed(224)                                   # args.py:43-46
    model = MMBiDAF(HIDDEN, E_TEXT, E_AUDIO, E_IMAGE, device, drop_prob=DROP, max_transcript_length=M).to(device)
    trainer = Trainer(model)
    host = make_batch(CFG3["batch"], CFG3["lt"], CFG3["la"], CFG3["li"], CFG3["t_dec"], E_TEXT, E_AUDIO, E_IMAGE,
                      seed=224 + rank).pin()
    resident = host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(steps):
            fn()
        end.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(end)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)        # max over ranks
        return float(ms.item()) / 1e3

    work = torch.cuda.Stream(device)                         # never the legacy stream: the step is graph-captured
    torch.cuda.set_stream(work)
    launches0 = ops.launch_count
    for _ in range(args.warmup):
        trainer.step(resident)
    launches_per_step = (ops.launch_count - launches0) // max(args.warmup, 1)
    graphed = not args.no_graph
    if graphed:
        # the step is ~1200 launches (launch-bound): capture forward+backward+all-reduce+clip+Adadelta once and
        # replay; every replay executes the same kernels (counted above) and draws fresh dropout masks
        trainer.capture(resident, warmup=1)
        run_resident = lambda: trainer.step_graphed()
    else:
        run_resident = lambda: trainer.step(resident)
    for _ in range(2):
        run_resident()
    with ClockSampler(local) as clocks:
        seconds = timed(run_resident, args.steps)
    launches = launches_per_step * args.steps
    videos = CFG3["batch"] * world * args.steps

    # End-to-end: every step's inputs start in pinned host memory and its loss is read back by the host.  The copy
    # of step k+1 is issued on a side stream while step k computes (a two-slot device staging area), the way an
    # input pipeline would feed the model; all of it happens inside the timed region.
    copy_stream = torch.cuda.Stream(device)
    slots = [host.to(device), host.to(device)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"k": 0}

    def prefetch(k):
        slot = slots[k & 1]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k & 1])                       # the step that last read this slot is done
            for dst, src in ((slot.text, host.text), (slot.audio, host.audio), (slot.images, host.images),
                             (slot.targets, host.targets)):
                dst.copy_(src, non_blocking=True)                        # H2D from pinned memory
            ready[k & 1].record(copy_stream)

    def e2e_step():
        k = state["k"]
        torch.cuda.current_stream().wait_event(ready[k & 1])
        loss = trainer.step_graphed(slots[k & 1]) if graphed else trainer.step(slots[k & 1])
        consumed[k & 1].record()
        prefetch(k + 1)
        state["k"] = k + 1
        return float(loss.item())                                        # D2H read of the step's loss

    for ev in consumed:
        ev.record()
    prefetch(0)
    sections = set(args.sections.split(","))
    e2e_seconds = float("nan")
    if "e2e" in sections:
        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        e2e_seconds = timed(e2e_step, args.steps)
    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        fast = args.precision == "fast"
        t_bidaf, algo = bidaf_microbench(device, 20, 5, ops.PREC_BF16 if fast else ops.PREC_FP32) \
            if "bidaf" in sections else (float("nan"), 1)
        achieved = algo / t_bidaf / 1e9
        t_bwd, algo_bwd = bidaf_microbench(device, 20, 5, ops.PREC_BF16, backward=True) \
            if "bidaf" in sections and fast else (float("nan"), 1)
        line = {"metric": METRIC, "value": round(videos / seconds, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(seconds / args.steps * 1e3, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 tcgen05 BiDAF + tf32 library GEMMs, fp32 accumulate / soft-max / LSTM" if fast else "f32",
                "data": "synthetic", "config": workload_config(world), "clocks": clocks.summary(),
                "e2e": {"value": round(videos / e2e_seconds, 2), "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes(),
                        "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_seconds / args.steps * 1e3, 3)},
                "gpu_launches": launches,
                "roofline": {"kernel": ("fused BiDAF forward, tcgen05 bf16 tier (bidaf_pack_kernel + bidaf_tc2_kernel: Q2C, "
                                        "C2QA and C2QB blocks in one launch, two per SM)" if fast else "fused BiDAF forward, fp32 tier (bidaf_pass_f32 x2)")
                                       + ", BASELINE config 2 (B=64, Lc=512, Lq=256, d=200)",
                             "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4),
                             "traffic": 151061248 if fast else None,      # dram read+write per forward, ncu (profiles/r01_bidaf_tc_ncu.md)
                             "peak_source": peak_src,
                             "algorithmic_bytes": algo, "us_per_forward": round(t_bidaf * 1e6, 1),
                             # the fused backward (prep + P^T pass + two dS passes + reduce) of the same op, same method
                             "backward": {"us": round(t_bwd * 1e6, 1), "algorithmic_bytes": algo_bwd,
                                          "achieved": round(algo_bwd / t_bwd / 1e9, 1),
                                          "frac": round(algo_bwd / t_bwd / 1e9 / peak, 4)}}}
        if world == 1 and not args.no_cpu_baseline:
            vps, _, cores, sample = cpu_training_throughput(4, 3, 1)
            line["cpu_baseline"] = {"value": round(vps, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--precision", default="fast", choices=["fast", "fp32"],
                    help="fast: bf16 tensor-core BiDAF + TF32 GEMMs (rel<=2e-2); fp32: rel<=1e-5 tier")
    ap.add_argument("--sections", default="step,e2e,bidaf", help="profiling aid: which GPU sections to run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
