"""Public-API contract of FasterQwen3TTS, mirroring the reference's fake-based unit tests
(tests/test_voice_clone_prompt_api.py:57-204, tests/test_sample_rate.py) -- CPU only, no engine."""
import inspect
import types

import numpy as np
import pytest
import torch

from faster_qwen3_tts import FasterQwen3TTS


def _m(base=None):
    return FasterQwen3TTS(base or types.SimpleNamespace(model=types.SimpleNamespace()), object(), object(), device="cpu")


def test_sample_rate_inference():
    assert _m(types.SimpleNamespace(model=types.SimpleNamespace(speech_tokenizer=types.SimpleNamespace(sample_rate=24000)))).sample_rate == 24000
    assert _m(types.SimpleNamespace(sample_rate=22050)).sample_rate == 22050
    assert _m().sample_rate == 24000
    with pytest.raises(AttributeError, match="speech_tokenizer"):
        _ = _m().speech_tokenizer


def test_signatures_and_defaults():
    sig = inspect.signature(FasterQwen3TTS.generate_voice_clone_streaming)
    names = list(sig.parameters)
    assert names[:14] == ["self", "text", "language", "ref_audio", "ref_text", "max_new_tokens", "min_new_tokens",
                          "temperature", "top_k", "top_p", "do_sample", "repetition_penalty", "chunk_size", "xvec_only"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["chunk_size"] == 12 and d["non_streaming_mode"] is None and d["parity_mode"] is False
    sig = inspect.signature(FasterQwen3TTS.generate_voice_clone)
    assert list(sig.parameters)[:12] == ["self", "text", "language", "ref_audio", "ref_text", "max_new_tokens",
                                         "min_new_tokens", "temperature", "top_k", "top_p", "do_sample",
                                         "repetition_penalty"]
    sig = inspect.signature(FasterQwen3TTS.generate_custom_voice_streaming)
    assert list(sig.parameters)[:6] == ["self", "text", "speaker", "language", "instruct", "non_streaming_mode"]
    sig = inspect.signature(FasterQwen3TTS.generate_voice_design)
    assert list(sig.parameters)[:5] == ["self", "text", "instruct", "language", "non_streaming_mode"]
    sig = inspect.signature(FasterQwen3TTS.from_pretrained)
    assert list(sig.parameters)[:6] == ["model_name", "device", "dtype", "attn_implementation", "max_seq_len", "backend"]
    assert sig.parameters["max_seq_len"].default == 2048


def test_warmup_captures_once_and_alias():
    calls = []

    class G:
        def __init__(self, n):
            self.n = n

        def capture(self, **kw):
            calls.append((self.n, kw))

    m = FasterQwen3TTS(types.SimpleNamespace(model=types.SimpleNamespace()), G("p"), G("t"), device="cpu")
    m._warmup(42)
    m.warmup(7)
    assert calls == [("p", {"num_warmup": 3}), ("t", {"prefill_len": 42, "num_warmup": 3})]


def test_errors_match_reference():
    m = _m()
    with pytest.raises(NotImplementedError, match="ref_spk/ref_rvq"):
        m.generate_voice_clone("hi", "English", ref_audio="x.wav", ref_spk="a.spk")
    with pytest.raises(NotImplementedError):
        m.generate("hi")
    with pytest.raises(ValueError, match="Unsupported backend"):
        FasterQwen3TTS.from_pretrained("x", backend="nope")
    if not torch.cuda.is_available():
        with pytest.raises(ValueError, match="CUDA"):
            FasterQwen3TTS.from_pretrained("synthetic:0.6B", device="cpu")
    assert m._resolve_non_streaming_mode(None, default=True) is True
    assert m._resolve_non_streaming_mode(False, default=True) is False


def test_streaming_window_policy_matches_reference_arithmetic():
    """model.py:1052-1135 with a fake decoder that emits exactly 1920 samples per frame whose value encodes the
    frame index: every yielded chunk must contain exactly the new frames, for ICL (ref codes) and x-vector modes."""
    class Tok:
        sample_rate = 24000

        def decode(self, payload):
            c = payload["audio_codes"][0]  # [T,16]
            return [c[:, 0].float().repeat_interleave(1920)], 24000

    for ref in (None, torch.full((30, 16), -1, dtype=torch.long)):
        m = _m(types.SimpleNamespace(model=types.SimpleNamespace(speech_tokenizer=Tok())))
        frames = torch.arange(70).view(-1, 1).repeat(1, 16)
        chunks = [(frames[i:i + 8], {"chunk_steps": min(8, 70 - i)}) for i in range(0, 70, 8)]
        out = list(m._stream_audio(iter(chunks), Tok(), ref, 8))
        got = np.concatenate([a for a, _, _ in out])
        assert got.shape[0] == 70 * 1920
        assert np.array_equal(got[::1920], np.arange(70, dtype=np.float32))
        assert all(sr == 24000 for _, sr, _ in out)


def _cfg_model(H=8, ng=16):
    tc = types.SimpleNamespace(hidden_size=H, num_code_groups=ng)
    return _m(types.SimpleNamespace(model=types.SimpleNamespace(config=types.SimpleNamespace(talker_config=tc))))


def test_cached_reference_files_refused_like_the_reference_arrays_accepted():
    """.spk / .rvq PATHS (qwentts.cpp's formats) -> the reference's NotImplementedError (its own unit test,
    tests/test_voice_clone_prompt_api.py:116-134); the decoded form (speaker vector + codec frames as arrays) becomes the
    voice_clone_prompt dict of the prompt builder (in-memory half)."""
    m = _cfg_model()
    with pytest.raises(NotImplementedError, match="backend='ggml'"):
        m.generate_voice_clone(text="hello", language="English", ref_spk="speaker.spk")
    with pytest.raises(NotImplementedError, match="backend='ggml'"):
        next(m.generate_voice_clone_streaming(text="hello", language="English", ref_rvq="speaker.rvq"))
    assert m._cached_reference_prompt(None, None, {"x": 1}) == {"x": 1}
    emb = np.arange(8, dtype=np.float64)
    xv = m._cached_reference_prompt(emb, None, None)
    assert xv["x_vector_only_mode"] == [True] and xv["icl_mode"] == [False] and xv["ref_code"] == [None]
    assert xv["ref_spk_embedding"][0].dtype == torch.float32 and xv["ref_spk_embedding"][0].tolist() == emb.tolist()
    codes = np.arange(3 * 16, dtype=np.int32).reshape(3, 16)
    icl = m._cached_reference_prompt(emb, codes, None)
    assert icl["icl_mode"] == [True] and icl["x_vector_only_mode"] == [False]
    assert icl["ref_code"][0].dtype == torch.long and icl["ref_code"][0].shape == (3, 16)
    with pytest.raises(ValueError, match="ref_spk/ref_spk_emb is required"):
        m._cached_reference_prompt(None, codes, None)
    with pytest.raises(ValueError, match="must not be empty"):
        m._cached_reference_prompt(np.zeros(0), None, None)
    with pytest.raises(ValueError, match="the talker expects 8"):
        m._cached_reference_prompt(np.zeros(5), None, None)
    with pytest.raises(ValueError, match=r"\[T, 16\]"):
        m._cached_reference_prompt(emb, np.zeros((3, 4)), None)
    with pytest.raises(ValueError, match="not both"):
        m._cached_reference_prompt(emb, None, {"ref_spk_embedding": [emb]})
