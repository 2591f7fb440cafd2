"""Serving callers (the reference serialises requests behind a global lock,
examples/openai_server.py:71,181): `serving.ContinuousBatcher` on the REAL engine -- several voice-clone requests with
different texts / prompt lengths join one batched persistent kernel between chunks, their codec windows are decoded as
batches, and every request's audio must equal the audio of the same request served alone through the public streaming API
(greedy decoding: a row of a batched bf16 launch is bit-identical to the single-sequence kernel, tests/test_gpu_batch.py).
The ICL requests here have 184 prompt positions, i.e. they cross 192 cached keys while generating: from there the
single-sequence kernel would switch to its split-key attention (fp32 probabilities), which the batched kernel does not
have, and greedy tokens may flip at near-ties -- FQ3_ATTN_SPLIT=0 keeps both on the per-head path so that the comparison
isolates the serving layer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TEXTS = ["hello there general kenobi", "a much longer sentence that keeps going for a while so that the prompt lengths differ",
         "short one", "the quick brown fox jumps over the lazy dog"]


@pytest.mark.parametrize("codec_mode", ["window", "stateful"])
def test_continuous_batcher_serves_what_single_requests_produce(codec_mode, monkeypatch):
    monkeypatch.setenv("FQ3_ATTN_SPLIT", "0")
    from faster_qwen3_tts import FasterQwen3TTS
    from faster_qwen3_tts.serving import batcher_for_model, voice_clone_request
    m = FasterQwen3TTS.from_synthetic("tiny", dtype=torch.bfloat16, max_seq_len=512, seed=5, max_batch=4)
    m.streaming_codec = codec_mode
    m.predictor_graph.do_sample = False     # the fast path uses the predictor's own (baked) sampling switch
    gen = dict(max_new_tokens=21, min_new_tokens=21, do_sample=False)
    want = []
    for i, text in enumerate(TEXTS):
        parts = [pcm for pcm, sr, t in m.generate_voice_clone_streaming(text, "English", ref_audio="ref.wav", ref_text="ref words",
                                                                        chunk_size=8, xvec_only=(i % 2 == 0), **gen)]
        want.append(np.concatenate(parts))
        assert want[-1].shape[0] == 21 * 1920
    b = batcher_for_model(m, chunk_size=8)
    try:
        tickets = [b.submit(voice_clone_request(m, text, "English", "ref.wav", "ref words", xvec_only=(i % 2 == 0)), **gen)
                   for i, text in enumerate(TEXTS)]
        got = [t.audio() for t in tickets]
    finally:
        b.close()
    assert b.max_concurrent >= 2, "requests were served one at a time"
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape, i
        d = float(np.abs(g - w).max())
        print(f"request {i}: max|served - alone| = {d:.3e}")
        assert d == 0.0, i
