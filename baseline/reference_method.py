"""Reference-METHOD stand-in on the GPU: the comparison arm SURVEY.md section 8(d)(ii) prescribes when upstream ``qwen_tts``
is absent (it is: the reference's own package cannot be imported or installed offline).

NOT product code and NOT an oracle: ``bench.py`` times it beside the engine (``gpu_reference``) and one GPU test checks
that it generates the oracle's tokens.  Nothing under ``faster-qwen3-tts_b200/`` imports it.

What it restates is the reference's *method* for the decode path, over the same synthetic module tree the engine loads:

* ``RefTalkerGraph``    -- faster_qwen3_tts/talker_graph.py:21-214: a static, max_seq_len-long KV cache per layer
  (transformers ``StaticCache`` semantics: ``index_copy_`` of the new row, attention over the WHOLE buffer under an
  additive mask), a pre-built mask table with one row per position (``_build_attention_masks`` :71-95, rebuilt when the
  per-row left-pad key changes :177-190), ``position_ids = cache_position + rope_delta`` (:210-211), ONE
  ``torch.cuda.CUDAGraph`` of the single-token forward (:109-147) replayed per step after copying the input, the
  position and the mask row into static buffers (:198-214).
* ``RefPredictorGraph`` -- predictor_graph.py:23-214: the whole 15-pass loop (2-token prefill + 14 decodes, per-pass
  head, ``sample_logits`` with ``torch.multinomial``) captured as ONE CUDA graph over a 17-slot static cache.
* ``ref_generate_streaming`` -- streaming.py:57-188: eager prefill forward, ``prefill_kv`` (56 ``index_copy_``), then per
  frame the eager glue the reference runs between the two graph replays: ``token.item()``, embedding lookups, ``cat``
  + ``sum`` of 16 rows, trailing-text add, ``codec_head`` GEMV, ``stack`` + ``unique`` repetition penalty, top-k
  sampling, ``clone`` of the hidden state.

Layer arithmetic is the HF eager Qwen3 decoder block (the same arithmetic the oracle restates); sampling is the
product's line-by-line mirror of the reference's ``sampling.py``.  Weights are whatever the module tree holds
(synthetic in this image) -- the label in bench.py says so.
"""
from __future__ import annotations

import time
from typing import Generator, Optional, Tuple

import torch
import torch.nn.functional as F


def _rot(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def _sample(logits, *, temperature, top_k, top_p, do_sample, suppress_mask=None, suppress_tokens=None):
    """sampling.py:32-66 (graph-capturable: no host sync)."""
    logits = logits.clone()
    if suppress_mask is not None:
        logits = logits.masked_fill(suppress_mask, float("-inf"))
    if suppress_tokens:
        logits[..., suppress_tokens] = float("-inf")
    if not do_sample:
        return torch.argmax(logits, dim=-1)
    logits = logits / temperature
    if top_k > 0:
        tv, _ = torch.topk(logits, min(top_k, logits.size(-1)))
        logits = torch.where(logits < tv[..., -1:], torch.full_like(logits, float("-inf")), logits)
    if top_p < 1.0:
        sl, si = torch.sort(logits, descending=True, dim=-1)
        cum = torch.cumsum(F.softmax(sl, dim=-1), dim=-1)
        rem = cum > top_p
        rem[..., 1:] = rem[..., :-1].clone()
        rem[..., 0] = False
        logits = logits.masked_fill(rem.scatter(-1, si, rem), float("-inf"))
    return torch.multinomial(F.softmax(logits, dim=-1), 1).squeeze(-1)


def _penalty(logits, history, penalty):
    """sampling.py:10-29."""
    if penalty == 1.0 or history.numel() == 0:
        return logits
    uniq = history.reshape(-1).unique()
    t = logits[..., uniq]
    logits[..., uniq] = torch.where(t > 0, t / penalty, t * penalty)
    return logits


class _StaticStack:
    """One decoder stack over a static KV cache: forward(x [1,T,H], cache_position [T], cos/sin [1,T,128], mask)."""

    def __init__(self, stack, cfg, max_len: int, device, dtype):
        self.layers, self.norm, self.c = stack.layers, stack.norm, cfg
        nkv = cfg.num_key_value_heads
        self.k = [torch.zeros(1, nkv, max_len, 128, dtype=dtype, device=device) for _ in self.layers]
        self.v = [torch.zeros(1, nkv, max_len, 128, dtype=dtype, device=device) for _ in self.layers]
        self.max_len = max_len
        theta = getattr(cfg, "rope_theta", 1_000_000.0)
        self.inv_freq = (1.0 / (theta ** (torch.arange(0, 128, 2, dtype=torch.int64).to(torch.float32) / 128))).to(device)

    def reset(self):
        for k, v in zip(self.k, self.v):
            k.zero_()
            v.zero_()

    def rope(self, position_ids: torch.Tensor, dtype):
        """position_ids float [T] -> cos, sin [1,1,T,128] (HF rotary module, computed per call like upstream)."""
        fr = position_ids.to(torch.float32)[:, None] * self.inv_freq[None, :]
        emb = torch.cat((fr, fr), dim=-1)
        return emb.cos().to(dtype)[None, None], emb.sin().to(dtype)[None, None]

    def forward(self, x, cache_position, cos, sin, mask):
        c = self.c
        nH, nKV = c.num_attention_heads, c.num_key_value_heads
        T = x.shape[1]
        for li, l in enumerate(self.layers):
            a = l.self_attn
            h = l.input_layernorm(x)
            q = a.q_norm(a.q_proj(h).view(1, T, nH, 128)).transpose(1, 2)
            k = a.k_norm(a.k_proj(h).view(1, T, nKV, 128)).transpose(1, 2)
            v = a.v_proj(h).view(1, T, nKV, 128).transpose(1, 2)
            q = q * cos + _rot(q) * sin
            k = k * cos + _rot(k) * sin
            self.k[li].index_copy_(2, cache_position, k)     # StaticCache.update
            self.v[li].index_copy_(2, cache_position, v)
            rep = nH // nKV
            kk = self.k[li][:, :, None].expand(1, nKV, rep, self.max_len, 128).reshape(1, nH, self.max_len, 128)
            vv = self.v[li][:, :, None].expand(1, nKV, rep, self.max_len, 128).reshape(1, nH, self.max_len, 128)
            att = torch.matmul(q, kk.transpose(2, 3)) * (128 ** -0.5) + mask
            att = F.softmax(att, dim=-1, dtype=torch.float32).to(q.dtype)
            o = torch.matmul(att, vv).transpose(1, 2).reshape(1, T, nH * 128)
            x = x + a.o_proj(o)
            h = l.post_attention_layernorm(x)
            x = x + l.mlp.down_proj(F.silu(l.mlp.gate_proj(h)) * l.mlp.up_proj(h))
        return self.norm(x)


def _causal_row(pos: int, n: int, max_len: int, n_left_pad: int, dtype, device):
    """[1,1,n,max_len] additive mask for queries at cache slots pos..pos+n-1."""
    kpos = torch.arange(max_len, device=device)
    qpos = pos + torch.arange(n, device=device)
    ok = (kpos[None, :] <= qpos[:, None]) & (kpos[None, :] >= n_left_pad)
    m = torch.zeros(n, max_len, dtype=dtype, device=device)
    m.masked_fill_(~ok, torch.finfo(dtype).min)
    return m[None, None]


class RefTalkerGraph:
    def __init__(self, talker_model, talker_config, device="cuda", dtype=torch.bfloat16, max_seq_len=2048):
        self.device, self.dtype, self.max_seq_len = device, dtype, max_seq_len
        self.hidden_size = talker_config.hidden_size
        self.num_layers = talker_config.num_hidden_layers
        self.stack = _StaticStack(talker_model, talker_config, max_seq_len, device, dtype)
        self.input_buf = torch.zeros(1, 1, self.hidden_size, dtype=dtype, device=device)
        self.output_buf = torch.zeros(1, 1, self.hidden_size, dtype=dtype, device=device)
        self.cache_position = torch.zeros(1, dtype=torch.long, device=device)
        self.rope_deltas = torch.zeros(1, dtype=torch.float32, device=device)
        self.position_ids = torch.zeros(1, dtype=torch.float32, device=device)
        self.attn_mask = None
        self.attn_mask_table = None
        self._mask_key = "unset"
        self.graph = None

    def _build_attention_masks(self, n_left_pad: int = 0):
        self.attn_mask_table = [_causal_row(i, 1, self.max_seq_len, n_left_pad, self.dtype, self.device)
                                for i in range(self.max_seq_len)]
        if self.attn_mask is None:
            self.attn_mask = self.attn_mask_table[0].clone()
        else:
            self.attn_mask.copy_(self.attn_mask_table[0])

    def _decode_step(self):
        cos, sin = self.stack.rope(self.position_ids, self.dtype)
        out = self.stack.forward(self.input_buf, self.cache_position, cos, sin, self.attn_mask)
        self.output_buf.copy_(out)

    @torch.inference_mode()
    def capture(self, prefill_len=100, num_warmup=3):
        self._build_attention_masks(0)
        self._mask_key = (0,)
        self.cache_position[0] = prefill_len
        self.attn_mask.copy_(self.attn_mask_table[prefill_len])
        for _ in range(num_warmup):
            self._decode_step()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._decode_step()
            torch.cuda.synchronize()
            with torch.cuda.graph(self.graph):
                self._decode_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()

    def prefill_kv(self, past_key_values) -> int:
        self.stack.reset()
        seq_len = 0
        for li in range(self.num_layers):
            k, v = past_key_values[li]
            seq_len = k.shape[2]
            if seq_len > self.max_seq_len:
                raise RuntimeError(f"Input is too long: prefill has {seq_len} tokens but max_seq_len={self.max_seq_len}. "
                                   "Use shorter text or shorter reference audio.")
            pos = torch.arange(seq_len, device=self.device)
            self.stack.k[li].index_copy_(2, pos, k)
            self.stack.v[li].index_copy_(2, pos, v)
        return seq_len

    def set_generation_state(self, attention_mask, rope_deltas):
        pads = (0,)
        if attention_mask is not None:
            pads = tuple((attention_mask == 0).sum(dim=-1).tolist())
        if pads != self._mask_key:
            self._build_attention_masks(int(pads[0]))
            self._mask_key = pads
        if rope_deltas is None:
            self.rope_deltas.zero_()
        else:
            self.rope_deltas.copy_(rope_deltas.reshape(-1)[:1].to(self.rope_deltas))

    @torch.inference_mode()
    def run(self, input_embeds: torch.Tensor, position: int) -> torch.Tensor:
        self.input_buf.copy_(input_embeds)
        self.cache_position[0] = position
        self.attn_mask.copy_(self.attn_mask_table[position])
        self.position_ids.copy_(self.rope_deltas + self.cache_position[0].to(self.rope_deltas.dtype))
        self.graph.replay()
        return self.output_buf


class RefPredictorGraph:
    def __init__(self, code_predictor, pred_config, talker_hidden_size, device="cuda", dtype=torch.bfloat16,
                 do_sample=True, top_k=50, top_p=1.0, temperature=0.9):
        self.device, self.dtype = device, dtype
        self.num_codebooks = getattr(pred_config, "num_code_groups", 16) - 1
        self.max_seq = 2 + self.num_codebooks
        self.do_sample, self.top_k, self.top_p, self.temperature = do_sample, top_k, top_p, temperature
        cp = code_predictor
        self.small_to_mtp, self.lm_heads, self.codec_embeds = cp.small_to_mtp_projection, cp.lm_head, cp.model.codec_embedding
        self.stack = _StaticStack(cp.model, pred_config, self.max_seq, device, dtype)
        self.prefill_cache_pos = torch.arange(2, device=device)
        self.decode_cache_positions = [torch.tensor([2 + i], device=device) for i in range(self.num_codebooks - 1)]
        self.input_buf = torch.zeros(1, 2, talker_hidden_size, dtype=dtype, device=device)
        self.output_tokens = torch.zeros(self.num_codebooks, dtype=torch.long, device=device)
        self.prefill_attn = _causal_row(0, 2, self.max_seq, 0, dtype, device)
        self.decode_attn = [_causal_row(2 + i, 1, self.max_seq, 0, dtype, device) for i in range(self.num_codebooks - 1)]
        self.graph = None

    def _full_loop(self):
        kw = dict(temperature=self.temperature, top_k=self.top_k, top_p=self.top_p, do_sample=self.do_sample)
        h = self.small_to_mtp(self.input_buf)
        cos, sin = self.stack.rope(self.prefill_cache_pos, self.dtype)
        h = self.stack.forward(h, self.prefill_cache_pos, cos, sin, self.prefill_attn)
        tok = _sample(self.lm_heads[0](h[:, -1:, :])[:, 0, :], **kw)
        self.output_tokens[0] = tok[0]
        for cb in range(1, self.num_codebooks):
            emb = self.small_to_mtp(self.codec_embeds[cb - 1](tok.unsqueeze(0)))
            pos = self.decode_cache_positions[cb - 1]
            cos, sin = self.stack.rope(pos, self.dtype)
            h = self.stack.forward(emb, pos, cos, sin, self.decode_attn[cb - 1])
            tok = _sample(self.lm_heads[cb](h[:, -1:, :])[:, 0, :], **kw)
            self.output_tokens[cb] = tok[0]
        return self.output_tokens

    @torch.inference_mode()
    def capture(self, num_warmup=3):
        for _ in range(num_warmup):
            self.stack.reset()
            self._full_loop()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.graph = torch.cuda.CUDAGraph()
            self.stack.reset()
            self._full_loop()
            torch.cuda.synchronize()
            self.stack.reset()
            with torch.cuda.graph(self.graph):
                self._full_loop()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()

    @torch.inference_mode()
    def run(self, pred_input: torch.Tensor) -> torch.Tensor:
        self.input_buf.copy_(pred_input)
        self.stack.reset()
        self.graph.replay()
        return self.output_tokens.clone()


@torch.inference_mode()
def ref_generate_streaming(talker, tie, tam, tth, tpe, config, predictor_graph, talker_graph, max_new_tokens=2048,
                           min_new_tokens=2, temperature=0.9, top_k=50, top_p=1.0, do_sample=True,
                           repetition_penalty=1.05, chunk_size=12) -> Generator[Tuple[torch.Tensor, dict], None, None]:
    """streaming.py:19-188, statement for statement (graph objects above instead of the engine)."""
    eos_id, n_groups, V = config.codec_eos_token_id, config.num_code_groups, config.vocab_size
    device = tie.device
    suppress_mask = torch.zeros(V, dtype=torch.bool, device=device)
    suppress_mask[max(0, V - 1024):] = True
    suppress_mask[eos_id] = False
    embed_cb0 = talker.get_input_embeddings()
    embeds_rest = talker.code_predictor.get_input_embeddings()
    head = talker.codec_head
    kw = dict(temperature=temperature, top_k=top_k, top_p=top_p, do_sample=do_sample, suppress_mask=suppress_mask)
    t0 = time.time()
    out = talker.forward(inputs_embeds=tie, attention_mask=tam, use_cache=True, output_hidden_states=True,
                         return_dict=True, trailing_text_hidden=tth, tts_pad_embed=tpe, generation_step=None,
                         past_hidden=None, past_key_values=None)
    past_hidden, gen_step = out.past_hidden, out.generation_step
    token = _sample(out.logits[:, -1, :], suppress_tokens=[eos_id] if min_new_tokens > 0 else None, **kw)
    prefill_len = talker_graph.prefill_kv(out.past_key_values)
    talker_graph.set_generation_state(tam, getattr(talker, "rope_deltas", None))
    torch.cuda.synchronize()
    t_prefill = time.time() - t0
    buf, firsts, total, idx = [], [], 0, 0
    t1 = time.time()
    for step in range(max_new_tokens):
        if token.item() == eos_id:
            break
        last = embed_cb0(token.unsqueeze(1))
        rest = predictor_graph.run(torch.cat((past_hidden, last), dim=1))
        buf.append(torch.cat([token.view(1), rest]).detach())
        firsts.append(token.detach())
        rows = [last] + [embeds_rest[i](rest[i].unsqueeze(0).unsqueeze(0)) for i in range(n_groups - 1)]
        x = torch.cat(rows, dim=1).sum(1, keepdim=True)
        x = x + (tth[:, gen_step].unsqueeze(1) if gen_step < tth.shape[1] else tpe)
        pos = prefill_len + step
        if pos >= talker_graph.max_seq_len - 1:
            break
        hidden = talker_graph.run(x, position=pos)
        logits = head(hidden[:, -1, :]).unsqueeze(0)
        if repetition_penalty != 1.0 and firsts:
            logits = _penalty(logits, torch.stack(firsts), repetition_penalty)
        token = _sample(logits.squeeze(0), suppress_tokens=[eos_id] if len(firsts) < min_new_tokens else None, **kw)
        past_hidden = hidden[:, -1:, :].clone()
        gen_step += 1
        if len(buf) >= chunk_size:
            torch.cuda.synchronize()
            total += len(buf)
            yield torch.stack(buf), {"chunk_index": idx, "chunk_steps": len(buf), "prefill_ms": t_prefill * 1000 if idx == 0 else 0,
                                     "decode_ms": (time.time() - t1) * 1000, "total_steps_so_far": total, "is_final": False}
            buf, idx, t1 = [], idx + 1, time.time()
    if buf:
        torch.cuda.synchronize()
        total += len(buf)
        yield torch.stack(buf), {"chunk_index": idx, "chunk_steps": len(buf), "prefill_ms": t_prefill * 1000 if idx == 0 else 0,
                                 "decode_ms": (time.time() - t1) * 1000, "total_steps_so_far": total, "is_final": True}


def build_reference_method(talker, cfg, device="cuda", dtype=torch.bfloat16, max_seq_len=2048, prefill_len=100):
    """-> (predictor_graph, talker_graph), both captured (model.py:239-255 `_warmup`)."""
    pg = RefPredictorGraph(talker.code_predictor, cfg.code_predictor_config, cfg.talker_config.hidden_size, device=device,
                           dtype=dtype)
    tg = RefTalkerGraph(talker.model, cfg.talker_config, device=device, dtype=dtype, max_seq_len=max_seq_len)
    pg.capture()
    tg.capture(prefill_len=min(prefill_len, max_seq_len - 1))
    return pg, tg
