"""Continuous batching for the serving callers.

The reference's servers run one request at a time behind a global lock (``examples/openai_server.py:71,181`` --
``_model_lock``; ``cli.py serve`` likewise): a second client waits for the first one's last chunk.  Here requests JOIN
and LEAVE the engine's request slots between chunks: one worker thread owns the GPU, admits queued requests into free
slots (prompt assembly + prefill), advances every active slot by one chunk with ONE persistent-kernel launch
(``BatchScheduler.step``), runs each slot's streaming codec window (model.py:1052-1135) and hands the PCM to the
request's own queue.  Wire helpers keep the reference's formats (``examples/openai_server.py:91-118``): 16-bit
little-endian PCM, streaming WAV header with unknown length.
"""
from __future__ import annotations

import io
import queue
import struct
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterator, List, Optional

import numpy as np

_DONE = object()


def to_pcm16(pcm: np.ndarray) -> bytes:
    """float32 [-1,1] -> raw 16-bit little-endian PCM (examples/openai_server.py:91-93)."""
    return np.clip(np.asarray(pcm, dtype=np.float32) * 32768, -32768, 32767).astype(np.int16).tobytes()


def wav_header(sample_rate: int, data_len: int = 0xFFFFFFFF) -> bytes:
    """WAV header; data_len = 0xFFFFFFFF for a stream of unknown length (examples/openai_server.py:96-112)."""
    n_channels, bits = 1, 16
    byte_rate = sample_rate * n_channels * bits // 8
    block_align = n_channels * bits // 8
    riff_size = 0xFFFFFFFF if data_len == 0xFFFFFFFF else 36 + data_len
    buf = io.BytesIO()
    buf.write(b"RIFF")
    buf.write(struct.pack("<I", riff_size))
    buf.write(b"WAVE")
    buf.write(b"fmt ")
    buf.write(struct.pack("<IHHIIHH", 16, 1, n_channels, sample_rate, byte_rate, block_align, bits))
    buf.write(b"data")
    buf.write(struct.pack("<I", data_len))
    return buf.getvalue()


def to_wav_bytes(pcm: np.ndarray, sample_rate: int) -> bytes:
    raw = to_pcm16(pcm)
    return wav_header(sample_rate, len(raw)) + raw


@dataclass
class Ticket:
    """Handle of one submitted request: iterate it for (pcm float32, sample_rate, timing) chunks."""
    rid: int
    prepare: Callable[[], tuple]          # -> (tie, tam, tth, tpe, ref_codes); runs on the worker thread
    gen_kwargs: dict
    out: "queue.Queue" = field(default_factory=queue.Queue)
    submitted_at: float = field(default_factory=time.time)
    first_chunk_at: Optional[float] = None
    frames: int = 0

    def __iter__(self) -> Iterator:
        while True:
            item = self.out.get()
            if item is _DONE:
                return
            if isinstance(item, BaseException):
                raise item
            yield item

    def audio(self) -> np.ndarray:
        parts = [c[0] for c in self]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=np.float32)


class ContinuousBatcher:
    """One worker thread, many clients.  ``scheduler_factory()`` -> object with ``has_capacity()``,
    ``submit(tie, tam, tth, tpe, tag=..., **gen) -> request``, ``step(n) -> [(request, codes)]`` and ``__len__`` (the
    ``BatchScheduler`` of batching.py); ``window_factory(ref_codes)`` -> object with ``push(codes) -> (pcm, sr)``
    (``model._StreamWindow``)."""

    def __init__(self, scheduler, window_factory: Callable, chunk_size: int = 8, idle_sleep: float = 0.002,
                 batch_decode: Optional[Callable] = None):
        self.sched, self.window_factory, self.chunk_size = scheduler, window_factory, chunk_size
        self.batch_decode = batch_decode   # (windows, code chunks) -> [(pcm, sr)]: one codec batch per window length
        self.idle_sleep = idle_sleep
        self.pending: "queue.Queue[Ticket]" = queue.Queue()
        self.live: Dict[int, tuple] = {}     # rid -> (ticket, window)
        self._rid = 0
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self.steps = 0
        self.max_concurrent = 0
        self._thread = threading.Thread(target=self._run, name="fq3-batcher", daemon=True)
        self._thread.start()

    # ---- client side ------------------------------------------------------------------------------------
    def submit(self, prepare: Callable[[], tuple], **gen_kwargs) -> Ticket:
        with self._lock:
            self._rid += 1
            t = Ticket(self._rid, prepare, gen_kwargs)
        self.pending.put(t)
        return t

    def close(self):
        self._stop.set()
        self._thread.join(timeout=10)

    # ---- worker -----------------------------------------------------------------------------------------
    def _admit(self):
        while self.sched.has_capacity():
            try:
                t = self.pending.get_nowait()
            except queue.Empty:
                return
            try:
                tie, tam, tth, tpe, ref_codes = t.prepare()
                self.sched.submit(tie, tam, tth, tpe, tag=t.rid, **t.gen_kwargs)
                self.live[t.rid] = (t, self.window_factory(ref_codes))
            except BaseException as ex:   # a bad request must not take the worker down
                t.out.put(ex)
                t.out.put(_DONE)

    def _run(self):
        while not self._stop.is_set():
            self._admit()
            if not len(self.sched):
                time.sleep(self.idle_sleep)
                continue
            self.max_concurrent = max(self.max_concurrent, len(self.sched))
            try:
                results = self.sched.step(self.chunk_size)
            except BaseException as ex:
                for rid, (t, _) in list(self.live.items()):
                    t.out.put(ex)
                    t.out.put(_DONE)
                self.live.clear()
                continue
            self.steps += 1
            live = [(rq, codes) for rq, codes in results if int(codes.shape[0])]
            # windows of equal length of all requests are decoded as one batch when the window objects support it
            pcm_of = {}
            if live and self.batch_decode is not None:
                for (rq, _), (pcm, sr) in zip(live, self.batch_decode([self.live[rq.tag][1] for rq, _ in live],
                                                                      [codes for _, codes in live])):
                    pcm_of[rq.tag] = (pcm, sr)
            for rq, codes in results:
                t, win = self.live[rq.tag]
                n = int(codes.shape[0])
                if n:
                    pcm, sr = pcm_of[rq.tag] if rq.tag in pcm_of else win.push(codes)
                    if t.first_chunk_at is None:
                        t.first_chunk_at = time.time()
                    t.frames += n
                    t.out.put((pcm, sr, {"chunk_steps": n, "total_steps_so_far": t.frames,
                                         "is_final": bool(rq.finished)}))
                if rq.finished:
                    t.out.put(_DONE)
                    del self.live[rq.tag]


def batcher_for_model(model, chunk_size: int = 8, to_host: bool = True) -> ContinuousBatcher:
    """ContinuousBatcher over a ``FasterQwen3TTS`` whose engine was created with ``max_batch`` > 1."""
    from .batching import BatchScheduler
    from .model import _StreamWindow, decode_windows_batched
    m = model.model.model
    sched = BatchScheduler(model.engine, m.talker, m.config.talker_config, model.predictor_graph, model.talker_graph)
    st = m.speech_tokenizer

    class _CodesOnly:
        def push(self, codes):
            return codes.cpu().numpy(), model.sample_rate

    def window(ref_codes):
        return model._make_window(st, ref_codes, chunk_size, to_host) if st is not None else _CodesOnly()

    bd = (lambda wins, chunks: decode_windows_batched(st, wins, chunks)) if st is not None else None
    return ContinuousBatcher(sched, window, chunk_size=chunk_size, batch_decode=bd)


def voice_clone_request(model, text: str, language: str, ref_audio, ref_text: str = "", xvec_only: bool = False,
                        non_streaming_mode: bool = False, append_silence: bool = True):
    """``prepare`` callable for ``ContinuousBatcher.submit``: the voice-clone prompt path of the public API
    (model.py:465-543) up to the embeddings."""
    def prepare():
        _, _, _, tie, tam, tth, tpe, ref_codes = model._prepare_generation(
            text, ref_audio=ref_audio, ref_text=ref_text, language=language, xvec_only=xvec_only,
            non_streaming_mode=non_streaming_mode, append_silence=append_silence)
        return tie, tam, tth, tpe, ref_codes
    return prepare
