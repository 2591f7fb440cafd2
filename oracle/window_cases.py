"""Streaming codec-window cases shared by oracle/make_golden.py (which runs the REFERENCE's
`generate_voice_clone_streaming` body, model.py:1019-1137, on them) and tests/test_window_policy_cpu.py (which runs
the product's `_stream_audio`).  TEST INFRASTRUCTURE -- never imported by the product."""
from __future__ import annotations

import torch

SR = 24000


class FakeTokenizer:
    """Deterministic stand-in for `speech_tokenizer.decode`: `spf` samples per frame (minus `short` samples at the
    end, to exercise the calibration / rounding arithmetic), every sample depends on the frame's codes AND on the
    previous frame's, so decoding with or without left context gives different audio."""

    def __init__(self, spf: int = 16, short: int = 0, as_numpy: bool = False):
        self.spf, self.short, self.as_numpy = spf, short, as_numpy
        self.calls = []

    def decode(self, payload):
        codes = payload["audio_codes"]
        assert codes.dim() == 3 and codes.shape[0] == 1 and codes.shape[2] == 16
        c = codes[0].double()
        T = c.shape[0]
        self.calls.append(T)
        key = c[:, 0] * 0.001 + c[:, 5] * 0.01
        prev = torch.cat([key[:1] * 0.0, key[:-1]])
        base = (key + 0.5 * prev)[:, None] + torch.arange(self.spf, dtype=torch.float64)[None] * 1e-4
        audio = base.reshape(-1)[: T * self.spf - self.short].float()
        return [audio.numpy() if self.as_numpy else audio], SR


def chunk_stream(sizes, seed):
    """what fast_generate_streaming yields: (codes [n,16] int64, timing dict)"""
    g = torch.Generator().manual_seed(seed)
    total = 0
    out = []
    for i, n in enumerate(sizes):
        total += n
        out.append((torch.randint(0, 256, (n, 16), generator=g),
                    dict(chunk_index=i, chunk_steps=n, decode_ms=1.0, total_steps_so_far=total,
                         is_final=i == len(sizes) - 1)))
    return out


# name -> (chunk_size, chunk sizes as generated, reference frames (ICL) or 0, tokenizer kwargs)
CASES = {
    "chunk8_xvec": (8, [8, 8, 8, 8, 8, 3], 0, {}),
    "chunk12_icl": (12, [12, 12, 12, 5], 30, {}),
    "chunk30_calibrates_at_chunk": (30, [30, 30, 7], 0, {}),
    "chunk1": (1, [1] * 31, 0, {"as_numpy": True}),
    "chunk8_icl_short_decoder": (8, [8, 8, 8, 8, 8, 8], 17, {"short": 5}),
    "chunk4_eos_early": (4, [4, 4, 2], 0, {}),
}


def ref_codes_for(n, seed=99):
    if not n:
        return None
    return torch.randint(0, 256, (n, 16), generator=torch.Generator().manual_seed(seed))


# non-streaming decode + proportional reference trim (model.py:914-938): name -> (generated frames or None, reference
# frames, tokenizer kwargs)
NONSTREAM_CASES = {
    "nonstream_icl": (37, 29, {}),
    "nonstream_icl_short_decoder_numpy": (11, 40, {"short": 7, "as_numpy": True}),
    "nonstream_xvec": (20, 0, {}),
    "nonstream_no_tokens": (None, 12, {}),
}


def generated_codes(n, seed=5):
    if n is None:
        return None
    return torch.randint(0, 256, (n, 16), generator=torch.Generator().manual_seed(seed + n))


TIMING = dict(prefill_ms=3.0, decode_s=0.5, steps=10, ms_per_step=50.0, steps_per_s=20.0)
