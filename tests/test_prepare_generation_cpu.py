"""The reference's request-preparation flow (model.py:295-581) on the synthetic model, CPU only: behaviours the
reference pins in tests/test_voice_clone_prompt_api.py:259-381 plus shapes of every mode."""
import logging
import types

import pytest
import torch

from oracle import prompt_cases as PC

from faster_qwen3_tts.model import FasterQwen3TTS  # noqa: E402


@pytest.fixture(scope="module")
def tts():
    base = PC.build_base(seed=1)
    dummy = types.SimpleNamespace(engine=None)
    m = FasterQwen3TTS(base, dummy, dummy, device="cpu", dtype=torch.float32, max_seq_len=512)
    m._warmed_up = True     # no graphs to capture on CPU
    return m


def _prep(tts, **kw):
    with torch.inference_mode():
        return tts._prepare_generation(text="Hello there, general.", language="English", **kw)


def test_xvec_and_icl_modes_from_reference_audio(tts):
    m, talker, cfg, tie, tam, tth, tpe, ref_codes = _prep(tts, ref_audio="voice_a.wav", ref_text="the reference", xvec_only=True)
    H = cfg.hidden_size
    assert tie.shape[0] == 1 and tie.shape[2] == H and tam.shape == tie.shape[:2] and int(tam.min()) == 1
    assert tpe.shape == (1, 1, H) and tth.shape[2] == H and ref_codes is None
    p_x = tie.shape[1]
    *_, tie2, tam2, tth2, tpe2, ref_codes2 = _prep(tts, ref_audio="voice_a.wav", ref_text="the reference", xvec_only=False)
    assert ref_codes2 is not None and ref_codes2.shape[1] == 16
    assert tie2.shape[1] == p_x - 1 + ref_codes2.shape[0] + 1      # ICL prompt spans codec_bos + every reference frame
    # cached per (audio, text, mode, silence)
    assert ("voice_a.wav", "the reference", False, True) in tts._voice_prompt_cache
    # non-streaming: the whole text is in the prompt, trailing stream is the single pad vector
    *_, tie3, tam3, tth3, tpe3, _ = _prep(tts, ref_audio="voice_a.wav", ref_text="the reference", xvec_only=True,
                                          non_streaming_mode=True)
    assert tth3.shape[1] == 1 and torch.equal(tth3, tpe3) and tie3.shape[1] > p_x


def test_precomputed_xvec_prompt_needs_no_prompt_extraction(tts):
    called = []
    orig = tts.model.create_voice_clone_prompt
    tts.model.create_voice_clone_prompt = lambda *a, **k: called.append(1) or orig(*a, **k)
    try:
        spk = torch.randn(tts.model.model.config.talker_config.hidden_size)
        out = _prep(tts, voice_clone_prompt=dict(ref_spk_embedding=[spk]), ref_text="ignored for x-vector prompts")
    finally:
        tts.model.create_voice_clone_prompt = orig
    assert not called and out[-1] is None


def test_instruct_with_xvec_only_warns(tts, caplog):
    with caplog.at_level(logging.WARNING):
        _prep(tts, ref_audio="voice_a.wav", xvec_only=True, instruct="whisper")
    assert any("experimental" in r.message for r in caplog.records)


def test_prompt_dict_validation_matches_reference_messages(tts):
    spk = torch.randn(tts.model.model.config.talker_config.hidden_size)
    code = torch.zeros(12, 16, dtype=torch.long)
    with pytest.raises(ValueError, match="missing required keys"):
        _prep(tts, voice_clone_prompt=dict(ref_code=[code]))
    with pytest.raises(ValueError, match="must be a list with length 1"):
        _prep(tts, voice_clone_prompt=dict(ref_spk_embedding=spk))
    with pytest.raises(ValueError, match="must be opposites"):
        _prep(tts, voice_clone_prompt=dict(ref_spk_embedding=[spk], x_vector_only_mode=[True], icl_mode=[True]))
    with pytest.raises(ValueError, match="ref_code must be None in x_vector_only mode"):
        _prep(tts, voice_clone_prompt=dict(ref_spk_embedding=[spk], x_vector_only_mode=[True], ref_code=[code]))
    with pytest.raises(ValueError, match="ref_code is required in ICL mode"):
        _prep(tts, voice_clone_prompt=dict(ref_spk_embedding=[spk], x_vector_only_mode=[False], ref_code=[None]))
    with pytest.raises(ValueError, match="ref_text is required"):
        _prep(tts, voice_clone_prompt=dict(ref_spk_embedding=[spk], x_vector_only_mode=[False], icl_mode=[True],
                                           ref_code=[code]))
    out = _prep(tts, ref_text="now with text", voice_clone_prompt=dict(
        ref_spk_embedding=[spk], x_vector_only_mode=[False], icl_mode=[True], ref_code=[code]))
    assert out[-1] is code
    with pytest.raises(ValueError, match="ref_audio is required"):
        _prep(tts)


def test_upstream_prompt_items_are_accepted(tts):
    items = tts.model.create_voice_clone_prompt(ref_audio="voice_b.wav", ref_text="item text")
    out = _prep(tts, voice_clone_prompt=items)               # item.ref_text wins over the (empty) ref_text argument
    assert out[-1] is items[0].ref_code
    with pytest.raises(ValueError, match="must have length 1"):
        _prep(tts, voice_clone_prompt=items * 2)


def test_custom_voice_and_design_requests(tts):
    with torch.inference_mode():
        m, talker, cfg, tie, tam, tth, tpe = tts._prepare_generation_custom("Good morning.", "English", "Aiden")
        assert tth.shape[1] == 1                             # custom voice defaults to non-streaming text
        *_, tie_d, _, _, _ = tts._prepare_generation_custom("Good morning.", None, None, instruct="a calm old voice")
        assert tie_d.shape[1] > tie.shape[1] - 1             # instruct turn in front, no speaker row
        with pytest.raises(NotImplementedError, match="Speaker zorg not implemented"):
            tts._prepare_generation_custom("x", "English", "zorg")


def test_wav_reader_without_soundfile(tts, tmp_path):
    import wave
    import numpy as np
    p = tmp_path / "r.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes((np.ones((1600, 2)) * 1000).astype("<i2").tobytes())
    audio, sr = tts._load_ref_audio_with_silence(p, silence_secs=0.5)
    assert sr == 16000 and audio.ndim == 1 and audio.shape[0] == 1600 + 8000 and float(audio[-1]) == 0.0


def test_custom_voice_and_design_validate_like_the_reference(tts, monkeypatch):
    """model.py:1155-1167,1343-1346: model type check, upstream language / speaker validators, 0.6B drops `instruct`;
    the request never reaches the engine here (fast_generate is a double that records what it was handed)."""
    import faster_qwen3_tts.generate as G
    seen = {}

    def fake_fast_generate(**kw):
        seen["P"] = kw["talker_input_embeds"].shape[1]
        return None, {}

    monkeypatch.setattr(G, "fast_generate", fake_fast_generate)
    with pytest.raises(ValueError, match="Unsupported speaker"):
        tts.generate_custom_voice("hi", "zorg", "English")
    with pytest.raises(ValueError, match="Unsupported language"):
        tts.generate_custom_voice("hi", "Aiden", "Klingon")
    with pytest.raises(ValueError, match="Unsupported language"):
        tts.generate_voice_design("hi", "a calm voice", "Klingon")
    audio, sr = tts.generate_custom_voice("hi there", "Aiden", "English", instruct="cheerful")
    assert sr == tts.sample_rate and audio[0].shape == (1,)            # "no tokens" -> one zero sample (model.py:1199-1201)
    p_with = seen["P"]
    tts.model.model.tts_model_size = "0b6"
    try:
        tts.generate_custom_voice("hi there", "Aiden", "English", instruct="cheerful")
    finally:
        del tts.model.model.tts_model_size
    assert seen["P"] < p_with                                          # the instruct turn was dropped
    tts.model.model.tts_model_type = "base"
    try:
        with pytest.raises(ValueError, match="does not support custom voice"):
            tts.generate_custom_voice("hi", "Aiden", "English")
        with pytest.raises(ValueError, match="does not support voice design"):
            tts.generate_voice_design("hi", "x", "English")
    finally:
        del tts.model.model.tts_model_type
