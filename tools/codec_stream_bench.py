#!/usr/bin/env python3
"""Stateful codec streams vs the window policy, timed with CUDA events (full decoder geometry):
open_stream, warm(174 reference frames), push(8) on one stream, push(8) on 32 streams in one call, against the window
decodes they replace (1x33, 1x182, 32x33).  python tools/codec_stream_bench.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "faster-qwen3-tts_b200")]
import torch  # noqa: E402

from faster_qwen3_tts.codec import build_codec  # noqa: E402

st = build_codec(dtype=torch.bfloat16, device="cuda", seed=1)
g = torch.Generator().manual_seed(0)


def codes(*shape):
    return torch.randint(0, 2048, (*shape, 16), generator=g).cuda()


def ev_time(fn, n=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3


out = {}
for name, c in (("window_1x33", codes(1, 33)), ("window_1x182", codes(1, 182)), ("window_32x33", codes(32, 33))):
    out[name] = ev_time(lambda: st.decode({"audio_codes": c}))
ref, ch = codes(174), codes(8)
t0 = time.perf_counter()
s = st.open_stream()
torch.cuda.synchronize()
out["open_stream_first_ms"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
s2 = st.open_stream()
torch.cuda.synchronize()
out["open_stream_ms"] = (time.perf_counter() - t0) * 1e3


def warm():
    s.reset()
    s.warm(ref)


out["warm_174"] = ev_time(warm, n=5)
out["push_1x8"] = ev_time(lambda: s.push(ch))
streams = [st.open_stream() for _ in range(32)]
c32 = codes(32, 8)
out["push_32x8"] = ev_time(lambda: st.push_streams(streams, c32))
c8 = codes(8, 8)
out["push_8x8"] = ev_time(lambda: st.push_streams(streams[:8], c8))
print(json.dumps({k: ([round(x, 3) for x in v] if isinstance(v, tuple) else round(v, 3)) for k, v in out.items()}))
print("(pairs are [device ms by CUDA events, host wall ms] per call)")
