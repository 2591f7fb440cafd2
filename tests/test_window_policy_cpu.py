"""Streaming codec-window policy against fixtures recorded from the REFERENCE's own
`generate_voice_clone_streaming` body (model.py:1052-1135: accumulated Phase-1 decode with the ICL reference frames
prepended, calibration at max(25, chunk_size) frames, then 25-frame left-context windows trimmed by the calibrated
samples-per-frame), executed with deterministic doubles (oracle/make_golden.py::gen_window)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import prompt_cases as PC
from oracle import window_cases as WC

from faster_qwen3_tts.model import FasterQwen3TTS  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "window.npz")


@pytest.fixture(scope="module")
def tts():
    dummy = types.SimpleNamespace(engine=None)
    return FasterQwen3TTS(PC.build_base(seed=0), dummy, dummy, device="cpu", dtype=torch.float32, max_seq_len=512)


@pytest.mark.parametrize("name", sorted(WC.CASES))
def test_window_policy_reproduces_reference_chunks(tts, name):
    gold = np.load(GOLD)
    chunk_size, sizes, n_ref, tk = WC.CASES[name]
    tok = WC.FakeTokenizer(**tk)
    chunks = WC.chunk_stream(sizes, seed=len(name))
    got = list(tts._stream_audio(iter(chunks), tok, WC.ref_codes_for(n_ref), chunk_size, to_host=True))
    assert [len(a) for a, _, _ in got] == gold[name + "_lens"].tolist()
    assert tok.calls == gold[name + "_decoded_T"].tolist()          # same decode calls: same windows, same cost
    audio = np.concatenate([np.asarray(a, dtype=np.float32) for a, _, _ in got])
    assert np.array_equal(audio, gold[name + "_audio"])              # sample for sample
    assert all(sr == WC.SR for _, sr, _ in got)
    assert [t["chunk_index"] for _, _, t in got] == list(range(len(sizes)))


@pytest.mark.parametrize("name", sorted(WC.NONSTREAM_CASES))
def test_non_streaming_decode_and_reference_trim(tts, name, monkeypatch):
    """generate_voice_clone (model.py:895-951): ICL reference frames are prepended for the decode and cut off
    proportionally afterwards; no tokens -> a single zero sample."""
    import faster_qwen3_tts.generate as G
    gold = np.load(GOLD)
    n_gen, n_ref, tk = WC.NONSTREAM_CASES[name]
    tok = WC.FakeTokenizer(**tk)
    ref_codes = WC.ref_codes_for(n_ref)
    monkeypatch.setattr(G, "fast_generate", lambda **kw: (WC.generated_codes(n_gen), dict(WC.TIMING)))
    monkeypatch.setattr(tts, "_prepare_generation",
                        lambda **kw: (types.SimpleNamespace(speech_tokenizer=tok), None, None, None, None, None, None, ref_codes))
    audio, sr = tts.generate_voice_clone("text", "English", ref_audio="x.wav")
    assert sr == WC.SR and tok.calls == gold[name + "_decoded_T"].tolist()
    assert np.array_equal(np.asarray(audio[0], dtype=np.float32), gold[name + "_audio"])
