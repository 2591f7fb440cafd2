"""Streaming generation: same signature and per-chunk timing keys as the reference's ``fast_generate_streaming``
(faster_qwen3_tts/streaming.py:19-188).  One persistent-kernel launch produces one chunk of ``chunk_size`` frames;
the only host synchronisation is the end-of-chunk result read (the reference synchronises there too,
streaming.py:158)."""
from __future__ import annotations

import time
from typing import Generator, Optional, Tuple

import torch

from .generate import _sync, begin_fused, shared_engine, stepwise_frames


def _timing(idx, n, t_prefill, dt, total, final):
    return {"chunk_index": idx, "chunk_steps": n, "prefill_ms": t_prefill * 1000 if idx == 0 else 0,
            "decode_ms": dt * 1000, "total_steps_so_far": total, "is_final": final}


@torch.inference_mode()
def fast_generate_streaming(
    talker,
    talker_input_embeds: torch.Tensor,
    attention_mask: torch.Tensor,
    trailing_text_hiddens: torch.Tensor,
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
    uniforms: Optional[torch.Tensor] = None,
) -> Generator[Tuple[torch.Tensor, dict], None, None]:
    """Yields (codes LongTensor[chunk_steps,16], timing); the last chunk may be short and has is_final=True
    only when it is a partial chunk, exactly like the reference."""
    device = talker_input_embeds.device
    skw = dict(max_new_tokens=max_new_tokens, min_new_tokens=min_new_tokens, temperature=temperature, top_k=top_k,
               top_p=top_p, do_sample=do_sample, repetition_penalty=repetition_penalty)
    engine = shared_engine(predictor_graph, talker_graph)
    t0 = time.time()
    total = idx = 0
    if engine is not None:
        begin_fused(engine, talker, talker_input_embeds, attention_mask, trailing_text_hiddens, tts_pad_embed, config,
                    predictor_graph, talker_graph, uniforms=uniforms, **skw)
        _sync(device)
        t_prefill = time.time() - t0
        t1 = time.time()
        while True:
            codes, res = engine.decode_chunk(chunk_size, slot=getattr(talker_graph, "slot", 0))
            n = res.frames_emitted
            if n:
                total += n
                tm = _timing(idx, n, t_prefill, time.time() - t1, total, n < chunk_size)
                if engine.time_kernels:
                    tm["kernel_ms"] = engine.last_kernel_ms
                yield codes.clone(), tm
                idx += 1
                t1 = time.time()
            if res.finished:
                return
    else:
        buf, t_prefill, t1 = [], 0.0, t0
        for kind, row in stepwise_frames(talker, talker_input_embeds, attention_mask, trailing_text_hiddens,
                                         tts_pad_embed, config, predictor_graph, talker_graph, **skw):
            if kind == "prefill_done":
                _sync(device)
                t_prefill = time.time() - t0
                t1 = time.time()
                continue
            buf.append(row)
            if len(buf) >= chunk_size:
                _sync(device)
                total += len(buf)
                yield torch.stack(buf), _timing(idx, len(buf), t_prefill, time.time() - t1, total, False)
                buf, idx, t1 = [], idx + 1, time.time()
        if buf:
            _sync(device)
            total += len(buf)
            yield torch.stack(buf), _timing(idx, len(buf), t_prefill, time.time() - t1, total, True)
