#!/usr/bin/env python3
"""Micro-benchmark of the batched persistent decode kernel: ms per frame-step and aggregate frames/s for B slots
(decode only, no codec).  python tools/batch_bench.py [--size 1.7B] [--prompt 40] [--frames 32] [--batches 1,2,4,8,16,32]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-qwen3-tts_b200"))
import torch  # noqa: E402

from faster_qwen3_tts import synthetic  # noqa: E402
from faster_qwen3_tts.batching import BatchScheduler  # noqa: E402
from faster_qwen3_tts.model import FasterQwen3TTS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="1.7B")
ap.add_argument("--prompt", type=int, default=40)
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--chunk", type=int, default=8)
ap.add_argument("--batches", default="1,2,4,8,16,32")
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--phases", action="store_true", help="clock64 phase accounting of CTA 0 over the last launch")
a = ap.parse_args()
Bs = [int(x) for x in a.batches.split(",")]
dt = torch.float32 if a.fp32 else torch.bfloat16
cfg = synthetic.make_config(a.size)
model = FasterQwen3TTS.from_synthetic(a.size, dtype=dt, with_codec=False, max_seq_len=2048, max_batch=max(Bs))
eng = model.engine
m = model.model.model
for B, mode in [(B, r) for B in Bs for r in range(a.repeat)]:

    def run():
        sched = BatchScheduler(eng, m.talker, m.config.talker_config, model.predictor_graph, model.talker_graph)
        for b in range(B):
            tie, tam, tth, tpe = synthetic.make_prompt(cfg, a.prompt, 4, seed=b, dtype=dt, device="cuda")
            sched.submit(tie, tam, tth, tpe, tag=b, max_new_tokens=a.frames, min_new_tokens=a.frames)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 0
        while len(sched):
            for rq, codes in sched.step(a.chunk):
                n += codes.shape[0]
        e1.record()
        e1.synchronize()
        return n, e0.elapsed_time(e1)
    run()
    if a.phases:
        eng.debug_enable(2)
        c0 = eng.probe_timestamps(6).tolist()
    n, ms = run()
    ph = None
    if a.phases:
        cyc = [b - x for b, x in zip(eng.probe_timestamps(6).tolist(), c0)]
        eng.debug_enable(0)
        names = ["other", "gemv", "barrier", "norm", "attn", "sample"]
        ph = {k: round(v / 1.965e3 / (n / B), 1) for k, v in zip(names, cyc)}   # us per frame-step at 1965 MHz (CTA 0)
    steps = n / B
    print(json.dumps({"B": B, "mode": mode, "size": a.size, "dtype": str(dt), "frames": n, "ms": round(ms, 3), "ms_per_frame_step": round(ms / steps, 4),
                      "agg_frames_per_s": round(n / ms * 1000, 1), "agg_rtf": round(n * 0.08 / (ms / 1000), 1), "phase_us_per_frame": ph}), flush=True)
