"""Offline stand-ins for the parts of upstream ``qwen_tts.Qwen3TTSModel`` that prompt assembly calls.

The reference's ``_prepare_generation`` (model.py:465-543) talks to the upstream wrapper through a handful of methods:
``_build_assistant_text`` / ``_build_ref_text`` / ``_build_instruct_text`` / ``_tokenize_texts`` (chat template +
tokenizer), ``create_voice_clone_prompt`` / ``_prompt_items_to_voice_clone_prompt`` (speaker encoder + codec encoder
on the reference audio) and, on the inner model, ``generate_speaker_prompt`` / ``generate_icl_prompt``.  None of
that exists offline, so the synthetic model answers the same calls deterministically:

* the tokenizer keeps the *layout* the reference slices by position -- 3 role tokens, the text, 5 closing tokens for
  an assistant turn (``input_id[:, :3]``, ``[:, 3:-5]``, model.py:674,705,733) and 3 + text + 2 for a reference turn
  (``ref_id[:, 3:-2]``, model.py:706) -- with word ids hashed into the synthetic text vocabulary;
* the voice-clone prompt is a seeded speaker vector plus seeded codec frames at 12.5 frames per second of the
  reference audio;
* ``generate_icl_prompt`` restates the public description of the ICL layout (BLOG.md:200-211: reference text and
  reference codec frames summed position by position, text continuing into the trailing stream or padded) -- an
  ANALOGUE of upstream code that is not available here; with a real checkpoint upstream's own method is called.
"""
from __future__ import annotations

import hashlib
import re
import types
import wave
from typing import List, Optional

import numpy as np
import torch

FRAME_RATE = 12.5
IM_START, IM_END, NL, ASSISTANT, USER = 1, 2, 3, 4, 5
FIRST_WORD_ID = 16


def _seed(*parts) -> int:
    h = hashlib.sha256("|".join(str(p) for p in parts).encode()).digest()
    return int.from_bytes(h[:4], "little")


class SyntheticInner(types.SimpleNamespace):
    """``base.model``: talker, config, speech_tokenizer + the two prompt hooks of the upstream inner model."""

    def generate_speaker_prompt(self, voice_clone_prompt) -> List[torch.Tensor]:
        dev, dt = self.talker.device, self.talker.codec_head.weight.dtype
        return [torch.as_tensor(e).to(device=dev, dtype=dt) for e in voice_clone_prompt["ref_spk_embedding"]]

    def generate_icl_prompt(self, text_id, ref_id, ref_code, tts_pad_embed, tts_eos_embed, non_streaming_mode):
        t = self.talker
        tc = self.config.talker_config
        text = t.text_projection(t.get_text_embeddings()(torch.cat([ref_id, text_id], dim=-1)))
        text = torch.cat([text, tts_eos_embed], dim=1)                                   # [1, Lt, H]
        frames = t.get_input_embeddings()(ref_code[:, 0])                                # codebook 0: talker table
        books = t.code_predictor.get_input_embeddings()
        for i in range(1, ref_code.shape[1]):
            frames = frames + books[i - 1](ref_code[:, i])                               # residual books: predictor tables
        bos = t.get_input_embeddings()(torch.tensor([tc.codec_bos_id], device=frames.device))
        codec = torch.cat([bos, frames], dim=0)[None]                                    # [1, Lc, H]
        lt, lc = text.shape[1], codec.shape[1]
        if non_streaming_mode:   # whole text first (on codec pad), then the reference frames on text pad
            pad = t.get_input_embeddings()(torch.full((1, lt), tc.codec_pad_id, device=frames.device))
            return torch.cat([text + pad, codec + tts_pad_embed], dim=1), tts_pad_embed
        if lt > lc:              # text outlasts the reference: the rest is fed step by step while generating
            return text[:, :lc] + codec, text[:, lc:]
        text = torch.cat([text, tts_pad_embed.expand(-1, lc - lt, -1)], dim=1)
        return text + codec, tts_pad_embed


class SyntheticOuter:
    """``base``: what ``Qwen3TTSModel`` is to the reference wrapper."""

    synthetic = True

    def __init__(self, inner: SyntheticInner):
        self.model = inner
        self.syn_cfg = None

    # ---- request validation (upstream raises ValueError for names it does not know) ---------------------------
    def _validate_languages(self, languages) -> None:
        known = self.model.config.talker_config.codec_language_id
        for lang in languages:
            if lang is not None and lang.lower() != "auto" and lang.lower() not in known:
                raise ValueError(f"Unsupported language: {lang}. Supported: {sorted(known)} or 'Auto'")

    def _validate_speakers(self, speakers) -> None:
        known = self.model.config.talker_config.spk_id
        for spk in speakers:
            if spk not in (None, "") and spk.lower() not in known:
                raise ValueError(f"Unsupported speaker: {spk}. Supported: {sorted(known)}")

    # ---- chat template + tokenizer -------------------------------------------------------------------------
    @staticmethod
    def _build_assistant_text(text: str) -> str:
        return f"<|im_start|>assistant\n{text}<|im_end|>\n<|im_start|>assistant\n"

    @staticmethod
    def _build_ref_text(text: str) -> str:
        return f"<|im_start|>assistant\n{text}<|im_end|>\n"

    @staticmethod
    def _build_instruct_text(text: str) -> str:
        return f"<|im_start|>user\n{text}<|im_end|>\n"

    def _word_ids(self, text: str) -> List[int]:
        vocab = self.model.talker.get_text_embeddings().num_embeddings
        out = []
        for w in re.findall(r"\w+|[^\w\s]", text, flags=re.UNICODE):
            pieces = [w] if len(w) <= 6 else [w[:6], "##" + w[6:]]
            out += [FIRST_WORD_ID + _seed("tok", p) % (vocab - FIRST_WORD_ID) for p in pieces]
        return out

    def _tokenize_texts(self, texts: List[str]) -> List[torch.Tensor]:
        special = {"<|im_start|>": IM_START, "<|im_end|>": IM_END, "\n": NL}
        res = []
        for s in texts:
            ids: List[int] = []
            for part in re.split(r"(<\|im_start\|>|<\|im_end\|>|\n)", s):
                if not part:
                    continue
                if part in special:
                    ids.append(special[part])
                elif part in ("assistant", "user") and ids and ids[-1] == IM_START:
                    ids.append(ASSISTANT if part == "assistant" else USER)
                else:
                    ids += self._word_ids(part)
            res.append(torch.tensor([ids], dtype=torch.long, device=self.model.talker.device))
        return res

    # ---- reference audio -> (speaker vector, codec frames) -------------------------------------------------
    @staticmethod
    def _audio_seconds(ref_audio) -> float:
        if isinstance(ref_audio, (tuple, list)) and len(ref_audio) == 2:
            a, sr = ref_audio
            return float(len(a)) / float(sr)
        try:
            with wave.open(str(ref_audio), "rb") as w:
                return w.getnframes() / float(w.getframerate())
        except Exception:
            return 174 / FRAME_RATE      # the bench's 13.9 s reference

    def create_voice_clone_prompt(self, ref_audio, ref_text: str = "", x_vector_only_mode: bool = False):
        tc = self.model.config.talker_config
        key = ref_audio if isinstance(ref_audio, str) else "array:%d" % int(self._audio_seconds(ref_audio) * 1000)
        g = torch.Generator().manual_seed(_seed("spk", key))
        spk = torch.randn(tc.hidden_size, generator=g)
        code = None
        if not x_vector_only_mode:
            n = max(1, int(round(self._audio_seconds(ref_audio) * FRAME_RATE)))
            pv = self.model.talker.code_predictor.model.config.vocab_size if hasattr(
                self.model.talker.code_predictor.model, "config") else self.model.talker.code_predictor.lm_head[0].out_features
            code = torch.randint(0, pv, (n, tc.num_code_groups), generator=g)
        return [types.SimpleNamespace(ref_code=code, ref_spk_embedding=spk, x_vector_only_mode=bool(x_vector_only_mode),
                                      icl_mode=not x_vector_only_mode, ref_text=ref_text or None)]

    @staticmethod
    def _prompt_items_to_voice_clone_prompt(items):
        return dict(ref_code=[it.ref_code for it in items], ref_spk_embedding=[it.ref_spk_embedding for it in items],
                    x_vector_only_mode=[bool(it.x_vector_only_mode) for it in items],
                    icl_mode=[bool(it.icl_mode) for it in items])
