"""Synthetic stand-in for the upstream ``qwen_tts.Qwen3TTSModel`` module tree.

The image has neither ``qwen-tts`` nor any checkpoint, so benchmarks, smoke tests and parity
tests run on *random-init weights of the real architecture*.  This module builds a tree with exactly the
attribute paths the reference (and ``weights.py``) touch on the real model:

    base.model.talker.model.layers[i].self_attn.{q,k,v,o}_proj / q_norm / k_norm, .mlp.*, .*_layernorm
    base.model.talker.model.norm / .codec_embedding       base.model.talker.codec_head
    base.model.talker.code_predictor.{model, lm_head[15], small_to_mtp_projection, get_input_embeddings()}
    base.model.talker.forward(inputs_embeds=..., attention_mask=..., ...)   (prefill, generate.py:107-118)
    base.model.config.talker_config          base.model.speech_tokenizer.decode({"audio_codes": ...})
    base.model.talker.{get_text_embeddings(), text_projection}, base.model.{generate_speaker_prompt,
    generate_icl_prompt}, base.{_tokenize_texts, _build_assistant_text, _build_ref_text, _build_instruct_text,
    create_voice_clone_prompt, _prompt_items_to_voice_clone_prompt}          (prompt assembly, model.py:295-805)

``SynTalker.forward`` is the variable-length prefill; like the reference's prefill (upstream HF eager forward) it
is plain library ops (torch / cuBLAS) -- the engine takes over from the first decode step.
"""
from __future__ import annotations

import types
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .weights import rope_tables

GEOMETRIES = {
    # name: (talker H, I, L), (predictor H, I, L), has_mtp   -- SURVEY.md App. A
    "1.7B": ((2048, 6144, 28), (1024, 3072, 5), True),
    "0.6B": ((1024, 3072, 28), (1024, 3072, 5), False),
    "tiny": ((512, 768, 3), (256, 512, 2), True),   # test geometry (4/2 heads, small vocabularies)
}


def make_config(size: str = "1.7B", *, talker_vocab: int = 3072, pred_vocab: int = 2048, eos: int = 2150,
                heads=(16, 8)) -> types.SimpleNamespace:
    (th, ti, tl), (ph, pi, pl), mtp = GEOMETRIES[size]
    if size == "tiny":
        talker_vocab, pred_vocab, eos, heads = 1280, 256, 300, (4, 2)

    def stack(h, i, l, v):
        return types.SimpleNamespace(hidden_size=h, intermediate_size=i, num_hidden_layers=l,
                                     num_attention_heads=heads[0], num_key_value_heads=heads[1], head_dim=128,
                                     rms_norm_eps=1e-6, rope_theta=1_000_000.0, vocab_size=v, sliding_window=None)

    t = stack(th, ti, tl, talker_vocab)
    t.codec_eos_token_id = eos
    t.num_code_groups = 16
    # special ids of the codec vocabulary used by prompt assembly (model.py:628-690); the real values come from the
    # checkpoint's config.json, these only have to be distinct, inside the special range and different from eos
    base = talker_vocab - 1024 if talker_vocab > 1024 else talker_vocab - 32
    t.codec_pad_id, t.codec_bos_id = eos - 2, eos - 1
    t.codec_think_id, t.codec_nothink_id, t.codec_think_bos_id, t.codec_think_eos_id = eos + 4, eos + 5, eos + 6, eos + 7
    t.codec_language_id = {"chinese": base + 2, "english": base + 3, "german": base + 4, "japanese": base + 5,
                           "cantonese": base + 6}
    t.spk_id = {"vivian": base + 20, "ryan": base + 21, "uncle_fu": base + 22, "aiden": base + 23, "serena": base + 24}
    t.spk_is_dialect = {"vivian": False, "ryan": False, "uncle_fu": "cantonese", "aiden": False, "serena": False}
    t.text_vocab_size = 512 if size == "tiny" else 4096
    t.text_hidden_size = 256 if size == "tiny" else 2048
    p = stack(ph, pi, pl, pred_vocab)
    p.num_code_groups = 16
    # text-side special ids (model.py:641-649)
    return types.SimpleNamespace(talker_config=t, code_predictor_config=p, has_mtp=mtp,
                                 tts_bos_token_id=8, tts_eos_token_id=9, tts_pad_token_id=10)


class RMSNorm(nn.Module):
    def __init__(self, n, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n), requires_grad=False)
        self.variance_epsilon = eps

    def forward(self, x):
        dt = x.dtype
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        return self.weight * xf.to(dt)


def _lin(i, o, bias=False):
    m = nn.Linear(i, o, bias=bias)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


class Attention(nn.Module):
    def __init__(self, c):
        super().__init__()
        qd, kd = c.num_attention_heads * 128, c.num_key_value_heads * 128
        self.q_proj, self.k_proj, self.v_proj = _lin(c.hidden_size, qd), _lin(c.hidden_size, kd), _lin(c.hidden_size, kd)
        self.o_proj = _lin(qd, c.hidden_size)
        self.q_norm, self.k_norm = RMSNorm(128, c.rms_norm_eps), RMSNorm(128, c.rms_norm_eps)


class MLP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.gate_proj, self.up_proj = _lin(c.hidden_size, c.intermediate_size), _lin(c.hidden_size, c.intermediate_size)
        self.down_proj = _lin(c.intermediate_size, c.hidden_size)


class Layer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self_attn, self.mlp = Attention(c), MLP(c)
        self.input_layernorm = RMSNorm(c.hidden_size, c.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(c.hidden_size, c.rms_norm_eps)


def _rot(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class Stack(nn.Module):
    """Decoder stack (``talker.model`` / ``code_predictor.model``)."""

    def __init__(self, c, embed_tables: int = 0, embed_dim: int = 0):
        super().__init__()
        self.config = c
        self.layers = nn.ModuleList([Layer(c) for _ in range(c.num_hidden_layers)])
        self.norm = RMSNorm(c.hidden_size, c.rms_norm_eps)
        if embed_tables == 1:
            self.codec_embedding = nn.Embedding(c.vocab_size, embed_dim)
        elif embed_tables > 1:
            self.codec_embedding = nn.ModuleList([nn.Embedding(c.vocab_size, embed_dim) for _ in range(embed_tables)])

    @torch.no_grad()
    def prefill(self, x: torch.Tensor, n_left_pad: int = 0):
        """x [P,H] -> (post-norm hidden [P,H], [(k,v)] per layer with k,v [1,n_kv,P,128]).  Eager semantics."""
        c = self.config
        P = x.shape[0]
        nH, nKV = c.num_attention_heads, c.num_key_value_heads
        pos = (torch.arange(P) - n_left_pad).clamp(min=0)
        cos, sin = rope_tables(c.rope_theta, max(int(pos.max()) + 1, 1))
        cos, sin = cos[pos].to(x.device, x.dtype)[None], sin[pos].to(x.device, x.dtype)[None]
        idx = torch.arange(P, device=x.device)
        allowed = (idx[None, :] <= idx[:, None]) & (idx[None, :] >= n_left_pad)
        mask = torch.zeros(P, P, dtype=x.dtype, device=x.device).masked_fill_(~allowed, torch.finfo(x.dtype).min)
        kvs = []
        for l in self.layers:
            a = l.self_attn
            h = l.input_layernorm(x)
            q = a.q_norm(a.q_proj(h).view(P, nH, 128)).transpose(0, 1)
            k = a.k_norm(a.k_proj(h).view(P, nKV, 128)).transpose(0, 1)
            v = a.v_proj(h).view(P, nKV, 128).transpose(0, 1)
            q = q * cos + _rot(q) * sin
            k = k * cos + _rot(k) * sin
            kvs.append((k[None].contiguous(), v[None].contiguous()))
            rep = nH // nKV
            kk = k[:, None].expand(nKV, rep, P, 128).reshape(nH, P, 128)
            vv = v[:, None].expand(nKV, rep, P, 128).reshape(nH, P, 128)
            att = torch.matmul(q, kk.transpose(1, 2)) * (128 ** -0.5) + mask[None]
            att = F.softmax(att, dim=-1, dtype=torch.float32).to(x.dtype)
            o = torch.matmul(att, vv).transpose(0, 1).reshape(P, nH * 128)
            x = x + a.o_proj(o)
            h = l.post_attention_layernorm(x)
            x = x + l.mlp.down_proj(F.silu(l.mlp.gate_proj(h)) * l.mlp.up_proj(h))
        return self.norm(x), kvs


class CodePredictor(nn.Module):
    def __init__(self, pc, talker_hidden, has_mtp, n_books=15):
        super().__init__()
        self.model = Stack(pc, embed_tables=n_books, embed_dim=talker_hidden)
        self.lm_head = nn.ModuleList([_lin(pc.hidden_size, pc.vocab_size) for _ in range(n_books)])
        self.small_to_mtp_projection = _lin(talker_hidden, pc.hidden_size, bias=True) if has_mtp else nn.Identity()

    def get_input_embeddings(self):
        return self.model.codec_embedding


class TextProjection(nn.Module):
    """text hidden -> talker hidden, two biased linears with SiLU in between (stand-in for upstream's resize MLP)"""

    def __init__(self, d_in, d_out):
        super().__init__()
        self.linear_fc1 = _lin(d_in, d_out, bias=True)
        self.linear_fc2 = _lin(d_out, d_out, bias=True)

    def forward(self, x):
        return self.linear_fc2(F.silu(self.linear_fc1(x)))


class Talker(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        tc = cfg.talker_config
        self.config = tc
        self.model = Stack(tc, embed_tables=1, embed_dim=tc.hidden_size)
        self.text_embedding = nn.Embedding(getattr(tc, "text_vocab_size", 512), getattr(tc, "text_hidden_size", 256))
        self.text_projection = TextProjection(self.text_embedding.embedding_dim, tc.hidden_size)
        self.codec_head = _lin(tc.hidden_size, tc.vocab_size)
        self.code_predictor = CodePredictor(cfg.code_predictor_config, tc.hidden_size, cfg.has_mtp,
                                            tc.num_code_groups - 1)
        self.rope_deltas = None

    def get_input_embeddings(self):
        return self.model.codec_embedding

    def get_text_embeddings(self):
        return self.text_embedding

    @property
    def device(self):
        return self.codec_head.weight.device

    @torch.no_grad()
    def forward(self, inputs_embeds=None, attention_mask=None, trailing_text_hidden=None, tts_pad_embed=None,
                generation_step=None, past_hidden=None, past_key_values=None, **kw):
        """Prefill only (generate.py:107-118 calls it with generation_step=None, past_key_values=None)."""
        if past_key_values is not None or generation_step is not None:
            raise NotImplementedError("synthetic talker implements the prefill call only")
        x = inputs_embeds[0]
        pad = 0
        if attention_mask is not None:
            pad = int((attention_mask[0] == 0).sum())
        hid, kvs = self.model.prefill(x, pad)
        logits = self.codec_head(hid[-1:])[None]  # [1,1,V]
        self.rope_deltas = torch.full((1, 1), -pad, device=x.device, dtype=torch.long) if pad else None
        return types.SimpleNamespace(logits=logits, past_key_values=kvs, past_hidden=hid[-1:][None].clone(),
                                     generation_step=0)


def build_base_model(cfg, state: Optional[Dict[str, torch.Tensor]] = None, *, seed: int = 0, dtype=torch.bfloat16,
                     device="cpu", speech_tokenizer=None, std: float = 0.02):
    """Module tree with weights from `state` (HF-style names, see oracle.make_weights) or seeded N(0, std^2)."""
    dev = torch.device(device)
    with torch.device(dev):
        talker = Talker(cfg)
    for prm in talker.parameters():
        prm.requires_grad_(False)
    if state is not None:
        sd = {k[len("talker."):]: v for k, v in state.items() if k.startswith("talker.")}
        missing, unexpected = talker.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        front = [k for k in missing if k.startswith(("text_embedding", "text_projection"))]
        assert len(front) == len(missing), [k for k in missing if k not in front]
        g = torch.Generator(device=dev).manual_seed(seed + 77)   # the oracle's weight sets have no text front end
        for name, p in talker.named_parameters():
            if name in front:
                p.normal_(0.0, 0.5 if "embedding" in name else std, generator=g)
    else:
        g = torch.Generator(device=dev).manual_seed(seed)
        for name, p in talker.named_parameters():
            if name.endswith("norm.weight") or "layernorm" in name:
                p.fill_(1.0)
            elif "codec_embedding" in name:
                p.normal_(0.0, 1.0, generator=g)
            elif "text_embedding" in name:
                p.normal_(0.0, 0.5, generator=g)
            elif "codec_head" in name or "lm_head" in name:
                p.normal_(0.0, std * 4, generator=g)
            else:
                p.normal_(0.0, std, generator=g)
    talker = talker.to(dtype=dtype)
    from .synthetic_frontend import SyntheticInner, SyntheticOuter
    icfg = types.SimpleNamespace(talker_config=cfg.talker_config,
                                 tts_bos_token_id=getattr(cfg, "tts_bos_token_id", 8),
                                 tts_eos_token_id=getattr(cfg, "tts_eos_token_id", 9),
                                 tts_pad_token_id=getattr(cfg, "tts_pad_token_id", 10))
    inner = SyntheticInner(talker=talker, config=icfg, speech_tokenizer=speech_tokenizer)
    return SyntheticOuter(inner)


def make_prompt(cfg, P: int, Tt: int, seed: int = 0, dtype=torch.bfloat16, device="cpu"):
    """Synthetic request: talker_input_embeds [1,P,H], attention_mask [1,P], trailing_text_hiddens [1,Tt,H],
    tts_pad_embed [1,1,H] (the tuple model.py:805 returns)."""
    g = torch.Generator().manual_seed(10_000 + seed)
    H = cfg.talker_config.hidden_size
    tie = torch.randn(P, H, generator=g).to(dtype)
    tth = torch.randn(Tt, H, generator=g).to(dtype)
    tpe = torch.randn(H, generator=g).to(dtype)
    return (tie[None].to(device), torch.ones(1, P, dtype=torch.long, device=device), tth[None].to(device),
            tpe[None, None].to(device))
