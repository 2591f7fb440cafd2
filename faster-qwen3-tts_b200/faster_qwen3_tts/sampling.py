"""Host-side sampling helpers with the reference's semantics (faster_qwen3_tts/sampling.py:10-66).

Used once per request for the first token (from the prefill logits) and by the step-wise compatibility loop.  The
per-frame draws of the fused path happen inside the persistent kernel with identical semantics.  ``u`` selects the
noise contract (inverse-CDF draw on a caller-supplied uniform); with ``u=None`` the draw is
``torch.multinomial`` like the reference."""
from __future__ import annotations

from typing import Iterable, Optional

import torch


def apply_repetition_penalty(logits: torch.Tensor, token_history: torch.Tensor, repetition_penalty: float) -> torch.Tensor:
    """HF-style penalty over the set of all previously generated ids (in place)."""
    if repetition_penalty == 1.0 or token_history.numel() == 0:
        return logits
    ids = torch.unique(token_history)
    picked = logits.index_select(-1, ids)
    scaled = torch.where(picked > 0, picked / repetition_penalty, picked * repetition_penalty)
    logits.index_copy_(-1, ids, scaled)
    return logits


def _filter(logits: torch.Tensor, temperature: float, top_k: int, top_p: float) -> torch.Tensor:
    neg = float("-inf")
    logits = logits / temperature
    if top_k > 0:
        kth = torch.topk(logits, min(top_k, logits.size(-1))).values[..., -1:]
        logits = logits.masked_fill(logits < kth, neg)
    if top_p < 1.0:
        srt, order = torch.sort(logits, descending=True)
        cum = torch.cumsum(torch.softmax(srt, dim=-1), dim=-1)
        drop = cum > top_p
        drop[..., 0] = False
        srt = srt.masked_fill(drop, neg)
        logits = torch.full_like(logits, neg).scatter(-1, order, srt)
    return logits


def sample_logits(logits: torch.Tensor, *, temperature: float, top_k: int, top_p: float, do_sample: bool,
                  suppress_mask: Optional[torch.Tensor] = None, suppress_tokens: Optional[Iterable[int]] = None,
                  u: Optional[float] = None) -> torch.Tensor:
    """suppress -> (argmax | temperature -> top-k (ties kept) -> top-p -> softmax -> draw)."""
    logits = logits.clone()
    if suppress_mask is not None:
        logits[..., suppress_mask] = float("-inf")
    if suppress_tokens:
        logits[..., list(suppress_tokens)] = float("-inf")
    if not do_sample:
        return torch.argmax(logits, dim=-1)
    probs = torch.softmax(_filter(logits, temperature, top_k, top_p), dim=-1)
    if u is None:
        return torch.multinomial(probs, 1).squeeze(-1)
    cdf = torch.cumsum(probs.float(), dim=-1)
    target = cdf[..., -1:] * float(u)
    idx = torch.searchsorted(cdf, target, right=True).clamp(max=probs.size(-1) - 1)
    return idx.squeeze(-1)
