"""Prompt assembly: token ids + voice prompt -> the four tensors the decode path starts from.

Restates ``FasterQwen3TTS._build_talker_inputs_local`` of the reference (model.py:583-805), i.e. SURVEY.md section
8(f) item 1, the step immediately upstream of prefill.  Same inputs, same outputs
``(talker_input_embeds [B,P,H] left-padded, attention_mask [B,P], trailing_text_hiddens [B,Tt,H] padded with the
tts-pad vector, tts_pad_embed [1,1,H])``, same exceptions; pinned against the reference's own function executed on
the synthetic module tree (``oracle/make_golden.py`` -> ``tests/golden/prompt.npz``, ``tests/test_prompt_cpu.py``).

Layout of one request (position by position the text-side and the codec-side embeddings are ADDED):

    text side : [instruct turn]  role(3)   pad .. pad  bos  | streaming: text[0]              | rest -> trailing stream
    codec side:                            think-block [speaker] pad | bos                    |
                                                                     | non-streaming: text.. eos / pad.. , pad / bos
                                                                     | ICL: upstream generate_icl_prompt(...)

The think-block is ``nothink, think_bos, think_eos`` (language "auto") or ``think, think_bos, <language>, think_eos``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch


def _language_id(tc, language: str, speaker: Optional[str]):
    """model.py:621-637: explicit language -> codec id; a dialect speaker overrides chinese / auto."""
    assert language is not None
    low = language.lower()
    if low == "auto":
        lang_id = None
    else:
        if low not in tc.codec_language_id:
            raise NotImplementedError(f"Language {language} not implemented")
        lang_id = tc.codec_language_id[low]
    if low in ("chinese", "auto") and speaker not in ("", None) and tc.spk_is_dialect[speaker.lower()]:
        lang_id = tc.codec_language_id[tc.spk_is_dialect[speaker.lower()]]
    return lang_id


def build_talker_inputs(m, input_ids: Sequence[torch.Tensor], ref_ids: Sequence[Optional[torch.Tensor]],
                        voice_clone_prompt: Optional[Dict[str, Any]], languages: Sequence[str],
                        speakers: Optional[Sequence[Optional[str]]], non_streaming_mode: bool,
                        instruct_ids: Optional[Sequence[Optional[torch.Tensor]]] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    talker, cfg, tc = m.talker, m.config, m.config.talker_config
    dev = talker.device
    n_req = len(input_ids)

    def text(ids: torch.Tensor) -> torch.Tensor:                 # [1,n] text ids -> [1,n,H]
        return talker.text_projection(talker.get_text_embeddings()(ids))

    def codec(ids: List[int], like: torch.Tensor) -> torch.Tensor:  # codec ids -> [1,n,H]
        return talker.get_input_embeddings()(torch.tensor([ids], device=dev, dtype=like.dtype))

    clone_vecs = m.generate_speaker_prompt(voice_clone_prompt) if voice_clone_prompt is not None else None
    speakers = list(speakers) if speakers is not None else [None] * n_req
    instruct_ids = list(instruct_ids) if instruct_ids is not None else [None] * n_req

    rows: List[torch.Tensor] = []
    trails: List[torch.Tensor] = []
    pad_e = None
    for i, (ids, language, speaker) in enumerate(zip(input_ids, languages, speakers)):
        # ---- who speaks (model.py:603-617)
        if clone_vecs is None:
            if speaker in ("", None):
                spk = None
            else:
                if speaker.lower() not in tc.spk_id:
                    raise NotImplementedError(f"Speaker {speaker} not implemented")
                spk = talker.get_input_embeddings()(torch.tensor(tc.spk_id[speaker.lower()], device=dev, dtype=ids.dtype))
        elif voice_clone_prompt["x_vector_only_mode"][i] or voice_clone_prompt["icl_mode"][i]:
            spk = clone_vecs[i]
        else:
            spk = None
        lang_id = _language_id(tc, language, speaker)

        bos_e, eos_e, pad_e = text(torch.tensor(
            [[cfg.tts_bos_token_id, cfg.tts_eos_token_id, cfg.tts_pad_token_id]], device=dev, dtype=ids.dtype)).chunk(3, dim=1)

        # ---- codec-side control row: think block, [speaker], pad, bos
        think = ([tc.codec_nothink_id, tc.codec_think_bos_id, tc.codec_think_eos_id] if lang_id is None else
                 [tc.codec_think_id, tc.codec_think_bos_id, lang_id, tc.codec_think_eos_id])
        pieces = [codec(think, ids)]
        if spk is not None:
            pieces.append(spk.view(1, 1, -1))
        pieces.append(codec([tc.codec_pad_id, tc.codec_bos_id], ids))
        control = torch.cat(pieces, dim=1)
        n_ctl = control.shape[1]
        # text side under all but the last control position: pad ... pad, bos
        under = torch.cat([pad_e.expand(-1, n_ctl - 2, -1), bos_e], dim=1) + control[:, :-1]

        seq = []
        if instruct_ids[i] is not None:
            seq.append(text(instruct_ids[i]))
        seq += [text(ids[:, :3]), under]

        icl = (voice_clone_prompt is not None and voice_clone_prompt.get("ref_code", None) is not None
               and voice_clone_prompt["icl_mode"][i])
        if icl:
            icl_embed, trail = m.generate_icl_prompt(
                text_id=ids[:, 3:-5], ref_id=ref_ids[i][:, 3:-2],
                ref_code=voice_clone_prompt["ref_code"][i].to(dev).clone(),
                tts_pad_embed=pad_e, tts_eos_embed=eos_e, non_streaming_mode=non_streaming_mode)
            seq.append(icl_embed)
        elif non_streaming_mode:
            body = ids[:, 3:-5]
            seq.append(torch.cat([text(body), eos_e], dim=1) + codec([tc.codec_pad_id] * (body.shape[1] + 1), ids))
            seq.append(pad_e + codec([tc.codec_bos_id], ids))
            trail = pad_e
        else:
            seq.append(text(ids[:, 3:4]) + control[:, -1:])
            trail = torch.cat([text(ids[:, 4:-5]), eos_e], dim=1)
        rows.append(torch.cat(seq, dim=1)[0])
        trails.append(trail[0])

    # ---- batch: prompts are LEFT padded with zeros (mask 0), trailing streams right padded with the pad vector
    hidden = rows[0].shape[-1]
    p_max = max(r.shape[0] for r in rows)
    embeds = rows[0].new_zeros(n_req, p_max, hidden)
    mask = torch.zeros(n_req, p_max, dtype=torch.long, device=embeds.device)
    for i, r in enumerate(rows):
        embeds[i, p_max - r.shape[0]:] = r
        mask[i, p_max - r.shape[0]:] = 1
    t_max = max(t.shape[0] for t in trails)
    trailing = pad_e.reshape(1, 1, hidden).expand(n_req, t_max, hidden).clone()
    for i, t in enumerate(trails):
        trailing[i, : t.shape[0]] = t
    return embeds, mask, trailing, pad_e
