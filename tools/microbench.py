#!/usr/bin/env python3
"""GPU micro-benchmarks of the persistent kernel's pieces (run under gpurun; prints one line per measurement)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "faster-qwen3-tts_b200")]
import torch
from faster_qwen3_tts import synthetic
from faster_qwen3_tts.model import FasterQwen3TTS
from faster_qwen3_tts.engine import SamplingParams

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n

size = sys.argv[1] if len(sys.argv) > 1 else "1.7B"
ctas = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = FasterQwen3TTS.from_synthetic(size, dtype=torch.bfloat16, with_codec=False, num_ctas=ctas)
eng = m.engine
H = eng.H
x = torch.randn(H, device="cuda").bfloat16()
tb, pb = eng.tape_bytes()
for pos in (16, 256, 1024, 2000):
    ms = timeit(lambda: eng.talker_step(x, pos))
    print(json.dumps({"what": "talker_step", "size": size, "ctas": eng.num_ctas, "pos": pos, "ms": ms, "GBps": tb / ms / 1e6}))
pi = torch.randn(2, H, device="cuda").bfloat16()
u = torch.rand(15, device="cuda")
for ds in (False, True):
    ms = timeit(lambda: eng.predictor_run(pi, SamplingParams(do_sample=ds), u))
    print(json.dumps({"what": "predictor_run", "do_sample": ds, "ms": ms, "GBps": pb / ms / 1e6}))
if hasattr(eng, "barrier_test"):
    for kind in (0, 1, 2, 3, 4):
        ms = timeit(lambda: eng.barrier_test(1000, kind), n=5)
        print(json.dumps({"what": "barrier", "kind": kind, "us_per_barrier": ms}))

# ---- phase timeline of CTA 0 (clock64 probes)
names = ["norm_in", "gemv_qkv", "B1", "attn", "B2", "gemv_o(+load)", "B3", "norm+gemv_gu", "B4", "gemv_dn(+load)", "B5"]
def timeline(label, fn, L):
    eng.debug_enable(2)
    fn(); torch.cuda.synchronize()
    ts = eng.probe_timestamps(12 * L).double()
    eng.debug_enable(0)
    ts = ts.view(L, 12)
    d = (ts[:, 1:] - ts[:, :-1]) / 1.965e3  # us at 1965 MHz
    segs = ["norm_in", "gemv_qkv", "B1", "attn", "B2", "load+gemv_o", "B3", "norm+gemv_gu", "B4", "load+gemv_dn", "B5"]
    avg = d[1:].mean(0) if L > 1 else d.mean(0)
    print(json.dumps({"what": "timeline_us", "label": label, **{k: round(float(v), 2) for k, v in zip(segs, avg)},
                      "layer_total": round(float(avg.sum()), 2)}))
timeline("talker pos256", lambda: eng.talker_step(x, 256), eng.talker_cfg["num_hidden_layers"])
timeline("talker pos2000", lambda: eng.talker_step(x, 2000), eng.talker_cfg["num_hidden_layers"])
timeline("predictor pass1", lambda: eng.predictor_run(pi, SamplingParams(do_sample=False), u), eng.pred_cfg["num_hidden_layers"])
