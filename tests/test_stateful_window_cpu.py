"""Host logic of the stateful streaming-codec path with a fake tokenizer: window selection
(`streaming_codec`), reference warm-up, per-chunk pushes, and the batched grouping of `decode_windows_batched`
(streams with the same number of new frames advance in ONE call; rows keep their order)."""
import types

import numpy as np
import pytest
import torch

from faster_qwen3_tts.model import FasterQwen3TTS, _StatefulWindow, _StreamWindow, decode_windows_batched


class FakeStream:
    def __init__(self, st):
        self.st, self.frames, self.warmed = st, 0, 0

    def push(self, codes):
        return self.st.push_streams([self], codes[None])[0]

    def close(self):
        pass

    def warm(self, codes):
        self.st.calls.append(("warm", int(codes.shape[0])))
        self.warmed += int(codes.shape[0])
        self.frames += int(codes.shape[0])


class FakeTokenizer:
    """decode(): 1920 samples per frame whose value is the frame's first code; streams: the same, frame by frame"""
    sample_rate = 24000
    supports_streams = True

    def __init__(self):
        self.calls = []

    def open_stream(self):
        self.calls.append(("open",))
        return FakeStream(self)

    def push_streams(self, streams, codes, want_pcm=True):
        self.calls.append(("push", len(streams), int(codes.shape[1])))
        out = []
        for s, c in zip(streams, codes):
            s.frames += int(c.shape[0])
            out.append(c[:, 0].float().repeat_interleave(1920))
        return out

    def decode(self, payload):
        c = payload["audio_codes"]
        self.calls.append(("decode", tuple(c.shape)))
        return [row[:, 0].float().repeat_interleave(1920) for row in c], self.sample_rate


def _owner(mode):
    o = types.SimpleNamespace(streaming_codec=mode, _to_numpy=FasterQwen3TTS._to_numpy)
    o._make_window = types.MethodType(FasterQwen3TTS._make_window, o)
    return o


def test_window_selection_and_validation():
    st = FakeTokenizer()
    assert isinstance(_owner("window")._make_window(st, None, 8), _StreamWindow)
    assert isinstance(_owner("stateful")._make_window(st, None, 8), _StatefulWindow)
    # a tokenizer without streams (upstream's) keeps the reference policy even when "stateful" is asked for
    plain = types.SimpleNamespace(decode=st.decode, sample_rate=24000)
    assert isinstance(_owner("stateful")._make_window(plain, None, 8), _StreamWindow)
    with pytest.raises(ValueError):
        _owner("bogus")._make_window(st, None, 8)


def test_stateful_stream_audio_warms_reference_once_and_pushes_each_chunk():
    st = FakeTokenizer()
    ref = torch.arange(5)[:, None].expand(5, 16)
    chunks = [torch.full((8, 16), 7), torch.full((8, 16), 9), torch.full((3, 16), 4)]
    out = list(FasterQwen3TTS._stream_audio(_owner("stateful"), ((c, {"i": i}) for i, c in enumerate(chunks)), st, ref, 8))
    assert st.calls == [("open",), ("warm", 5), ("push", 1, 8), ("push", 1, 8), ("push", 1, 3)]
    assert [a.shape[0] for a, _, _ in out] == [8 * 1920, 8 * 1920, 3 * 1920]
    assert isinstance(out[0][0], np.ndarray) and out[0][1] == 24000 and out[2][2] == {"i": 2}
    assert np.all(out[1][0] == 9.0)


def test_batched_stateful_windows_group_by_new_frames_and_keep_row_order():
    st = FakeTokenizer()
    o = _owner("stateful")
    wins = [o._make_window(st, None, 8, to_host=False) for _ in range(4)]
    st.calls.clear()
    chunks = [torch.full((8, 16), 1), torch.full((3, 16), 2), torch.full((8, 16), 3), torch.full((8, 16), 4)]
    res = decode_windows_batched(st, wins, chunks)
    assert sorted(st.calls) == [("push", 1, 3), ("push", 3, 8)]
    for (audio, sr), c in zip(res, chunks):
        assert sr == 24000 and audio.shape[0] == c.shape[0] * 1920 and float(audio[0]) == float(c[0, 0])


class FakeTokenizerWithTemplates(FakeTokenizer):
    """adds the per-voice template cache of SpeechTokenizer.reference_stream"""

    def __init__(self):
        super().__init__()
        self.templates = {}

    def reference_stream(self, ref_codes, create=True):
        key = tuple(int(x) for x in ref_codes[:, 0])
        if key not in self.templates:
            if not create:
                return None
            self.calls.append(("template_warm", len(key)))
            self.templates[key] = True
        s = FakeStream(self)
        s.frames = len(key)
        self.calls.append(("template_copy",))
        return s


def test_window_policy_phase1_new_voice_is_literal_first_then_streams_known_voice_streams_at_once(monkeypatch):
    """default ("window") policy with an ICL reference: a voice never seen takes the literal re-decode for its FIRST chunk
    (time to first audio) and warms its template when the second chunk arrives; a known voice streams from chunk one;
    Phase 2 is the reference's 25-frame window either way; the audio is the same as the literal policy's."""
    monkeypatch.delenv("FQ3_PHASE1_STREAM", raising=False)
    st = FakeTokenizerWithTemplates()
    ref = torch.arange(100, 110)[:, None].expand(10, 16)
    chunks = [torch.full((8, 16), v) for v in (1, 2, 3, 4, 5)]
    o = _owner("window")

    def run(tok):
        tok.calls.clear()
        return [a for a, _, _ in FasterQwen3TTS._stream_audio(o, ((c, {}) for c in chunks), tok, ref, 8, to_host=False)]

    new_voice = run(st)
    kinds = [c[0] for c in st.calls]
    assert kinds == ["decode", "template_warm", "template_copy", "warm", "push", "push", "push", "decode"], st.calls
    assert st.calls[0] == ("decode", (1, 18, 16)) and st.calls[3] == ("warm", 8)      # reference + 8 frames; catch-up of chunk 1
    assert st.calls[-1] == ("decode", (1, 33, 16))                                      # Phase 2: 25 context + 8 new
    known_voice = run(st)
    assert [c[0] for c in st.calls] == ["template_copy", "push", "push", "push", "push", "decode"], st.calls
    monkeypatch.setenv("FQ3_PHASE1_STREAM", "0")
    literal = run(st)
    assert [c[0] for c in st.calls] == ["decode"] * 5 and [c[1][1] for c in st.calls] == [18, 26, 34, 42, 33]
    for a, b, c in zip(new_voice, known_voice, literal):
        assert a.shape[0] == 8 * 1920 and torch.equal(a, b) and torch.equal(a, c)
