#!/usr/bin/env python3
"""Codec window decode: CUDA-event time of the front end (embedding mean, pre-transformer, ConvNeXt upsampling) and of the
waveform stack (C ABI) for a batch of windows.  python tools/codec_prof.py --batch 1,32 --frames 33 [--once]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "faster-qwen3-tts_b200"))
import torch  # noqa: E402

from faster_qwen3_tts.codec import Code2WavConfig, build_codec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", default="1,32")
ap.add_argument("--frames", type=int, default=33)
ap.add_argument("--once", action="store_true", help="one decode per batch size after warm-up (for ncu launch lists)")
a = ap.parse_args()
st = build_codec(Code2WavConfig(), seed=3, dtype=torch.bfloat16, device="cuda", backend="engine")
for B in [int(x) for x in a.batch.split(",")]:
    codes = torch.randint(0, 2048, (B, a.frames, 16), device="cuda")
    for _ in range(3):
        st.decode({"audio_codes": codes})
    torch.cuda.synchronize()
    if a.once:
        st.decode({"audio_codes": codes})
        torch.cuda.synchronize()
        continue
    n = 10
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(n):
        st.decode({"audio_codes": codes})
    e[1].record()
    c2 = codes.transpose(1, 2).contiguous()
    for _ in range(n):
        st._front_graphed(c2)
    e[2].record()
    e[2].synchronize()
    tot, front = e[0].elapsed_time(e[1]) / n, e[1].elapsed_time(e[2]) / n
    fl = st.flops(a.frames) * B
    print(json.dumps({"B": B, "T": a.frames, "decode_ms": round(tot, 3), "front_ms": round(front, 3), "stack_ms": round(tot - front, 3),
                      "stack_TFLOPs": round(fl / ((tot - front) / 1000) / 1e12, 1), "ms_per_window": round(tot / B, 3)}), flush=True)
