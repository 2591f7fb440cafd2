"""Batched decode over the engine's request slots (BASELINE config 4: concurrent requests per GPU).

The reference batches only the prompt: ``_build_talker_inputs_local`` left-pads a list of requests and returns a
per-row attention mask (faster_qwen3_tts/model.py:774-787), ``TalkerGraph.set_generation_state`` takes per-row pad
counts and rope deltas (talker_graph.py:177-187); its decode loop itself is batch-1 (``token.item()``,
generate.py:150).  Here every row of such a batch becomes one request *slot* of the engine and all active slots
advance together: one persistent-kernel launch per chunk in which the slots share every pass over the weight tape
(``fq3_decode_chunk(slots[], n_slots, ...)``).  Slots finish independently and can be re-used between chunks
(``BatchScheduler.submit`` while others are mid-stream = continuous batching for the serving callers,
examples/openai_server.py:71 serialises requests behind a lock instead).

Per row the arithmetic -- and therefore the codes -- is identical to running that request alone (tests/test_gpu_batch.py).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, Generator, List, Optional, Tuple

import torch

from .generate import _sync, begin_fused, shared_engine


@dataclass
class SlotRequest:
    slot: int
    tag: object
    max_new_tokens: int
    frames: int = 0
    finished: int = 0
    parts: List[torch.Tensor] = field(default_factory=list)


class BatchScheduler:
    """Owns the engine's slots: ``submit`` prefills one request into a free slot and latches it, ``step`` advances
    every active slot by up to ``n_frames`` frames with ONE launch and returns the new codes per request."""

    def __init__(self, engine, talker, config, predictor_graph, talker_graph):
        self.engine, self.talker, self.config = engine, talker, config
        self.pg, self.tg = predictor_graph, talker_graph
        self.free: List[int] = list(range(engine.max_batch))
        self.active: Dict[int, SlotRequest] = {}

    def __len__(self) -> int:
        return len(self.active)

    def has_capacity(self) -> bool:
        return bool(self.free)

    @torch.inference_mode()
    def submit(self, tie, tam, tth, tpe, *, tag=None, max_new_tokens: int = 2048, min_new_tokens: int = 2,
               temperature: float = 0.9, top_k: int = 50, top_p: float = 1.0, do_sample: bool = True,
               repetition_penalty: float = 1.05, uniforms: Optional[torch.Tensor] = None) -> SlotRequest:
        """One request: tie [1,P,H], tam [1,P] (zeros = left padding), tth [1,Tt,H], tpe [1,1,H]."""
        if not self.free:
            raise RuntimeError(f"all {self.engine.max_batch} request slots are busy")
        slot = self.free.pop(0)
        try:
            begin_fused(self.engine, self.talker, tie, tam, tth, tpe, self.config, self.pg, self.tg,
                        max_new_tokens=max_new_tokens, min_new_tokens=min_new_tokens, temperature=temperature,
                        top_k=top_k, top_p=top_p, do_sample=do_sample, repetition_penalty=repetition_penalty,
                        uniforms=uniforms, slot=slot)
        except Exception:
            self.free.insert(0, slot)
            raise
        rq = SlotRequest(slot=slot, tag=tag if tag is not None else slot, max_new_tokens=max_new_tokens)
        self.active[slot] = rq
        return rq

    @torch.inference_mode()
    def step(self, n_frames: int) -> List[Tuple[SlotRequest, torch.Tensor]]:
        """Advance all active slots; returns [(request, codes [n,16])] for every slot that was active (n may be 0 for a
        slot that stopped before emitting).  Finished slots are released."""
        if not self.active:
            return []
        slots = sorted(self.active)
        if len(slots) == 1:
            codes, res = self.engine.decode_chunk(n_frames, slot=slots[0])
            outs, ress = [codes], [res]
        else:
            buf, ress = self.engine.decode_chunk_batch(slots, n_frames)
            outs = [buf[j, : ress[j].frames_emitted] for j in range(len(slots))]
        done = []
        for s, c, r in zip(slots, outs, ress):
            rq = self.active[s]
            rq.frames += int(r.frames_emitted)
            rq.finished = int(r.finished)
            done.append((rq, c.clone()))
            if r.finished:
                del self.active[s]
                self.free.append(s)
        return done


def _rows(x: torch.Tensor, b: int) -> torch.Tensor:
    return x[b:b + 1]


@torch.inference_mode()
def fast_generate_streaming_batch(
    talker,
    talker_input_embeds: torch.Tensor,     # [B,P,H] left-padded (model.py:774-787)
    attention_mask: torch.Tensor,          # [B,P]
    trailing_text_hiddens: torch.Tensor,   # [B,Tt,H] (rows padded with tts_pad_embed, model.py:789-803)
    tts_pad_embed: torch.Tensor,
    config,
    predictor_graph,
    talker_graph,
    max_new_tokens: int = 2048,
    min_new_tokens: int = 2,
    temperature: float = 0.9,
    top_k: int = 50,
    top_p: float = 1.0,
    do_sample: bool = True,
    repetition_penalty: float = 1.05,
    chunk_size: int = 12,
    uniforms: Optional[torch.Tensor] = None,   # [B, max_new_tokens+1, 16]
) -> Generator[List[Tuple[int, torch.Tensor, dict]], None, None]:
    """Batched counterpart of ``fast_generate_streaming`` (streaming.py:19-188): every yielded item is the list of
    (row, codes [n,16], timing) of the rows that produced frames in that chunk; timing keys are the reference's."""
    engine = shared_engine(predictor_graph, talker_graph)
    if engine is None:
        raise RuntimeError("batched decode needs graph handles backed by one loaded fq3 engine")
    B = talker_input_embeds.shape[0]
    if B > engine.max_batch:
        raise ValueError(f"batch of {B} rows exceeds the engine's max_batch={engine.max_batch}")
    sched = BatchScheduler(engine, talker, config, predictor_graph, talker_graph)
    device = talker_input_embeds.device
    t0 = time.time()
    for b in range(B):
        sched.submit(_rows(talker_input_embeds, b), _rows(attention_mask, b), _rows(trailing_text_hiddens, b),
                     tts_pad_embed, tag=b, max_new_tokens=max_new_tokens, min_new_tokens=min_new_tokens,
                     temperature=temperature, top_k=top_k, top_p=top_p, do_sample=do_sample,
                     repetition_penalty=repetition_penalty, uniforms=None if uniforms is None else uniforms[b])
    _sync(device)
    t_prefill = time.time() - t0
    idx = 0
    totals = [0] * B
    while len(sched):
        t1 = time.time()
        out = sched.step(chunk_size)
        dt = time.time() - t1
        items = []
        for rq, codes in out:
            n = int(codes.shape[0])
            if not n:
                continue
            totals[rq.tag] += n
            tm = {"chunk_index": idx, "chunk_steps": n, "prefill_ms": t_prefill * 1000 if idx == 0 else 0,
                  "decode_ms": dt * 1000, "total_steps_so_far": totals[rq.tag], "is_final": n < chunk_size or rq.finished == 3}
            if engine.time_kernels:
                tm["kernel_ms"] = engine.last_kernel_ms
            items.append((rq.tag, codes, tm))
        if items:
            yield items
            idx += 1


@torch.inference_mode()
def fast_generate_batch(talker, talker_input_embeds, attention_mask, trailing_text_hiddens, tts_pad_embed, config,
                        predictor_graph, talker_graph, max_new_tokens: int = 2048, min_new_tokens: int = 2,
                        temperature: float = 0.9, top_k: int = 50, top_p: float = 1.0, do_sample: bool = True,
                        repetition_penalty: float = 1.05, uniforms: Optional[torch.Tensor] = None,
                        launch_frames: int = 64) -> Tuple[List[Optional[torch.Tensor]], dict]:
    """Batched counterpart of ``fast_generate`` (generate.py:16-215): (list of codes [steps_b,16] or None per row,
    timing with the reference's keys; ``steps`` is the total over rows)."""
    B = talker_input_embeds.shape[0]
    parts: List[List[torch.Tensor]] = [[] for _ in range(B)]
    t0 = time.time()
    prefill_ms = 0.0
    for items in fast_generate_streaming_batch(
            talker, talker_input_embeds, attention_mask, trailing_text_hiddens, tts_pad_embed, config, predictor_graph,
            talker_graph, max_new_tokens=max_new_tokens, min_new_tokens=min_new_tokens, temperature=temperature,
            top_k=top_k, top_p=top_p, do_sample=do_sample, repetition_penalty=repetition_penalty,
            chunk_size=launch_frames, uniforms=uniforms):
        for b, codes, tm in items:
            parts[b].append(codes)
            prefill_ms = max(prefill_ms, tm["prefill_ms"])
    _sync(talker_input_embeds.device)
    dt = time.time() - t0 - prefill_ms / 1000
    out = [torch.cat(p) if p else None for p in parts]
    n = sum(int(c.shape[0]) for c in out if c is not None)
    return out, {"prefill_ms": prefill_ms, "decode_s": dt, "steps": n, "ms_per_step": (dt / n * 1000) if n else 0,
                 "steps_per_s": (n / dt) if dt > 0 else 0}
