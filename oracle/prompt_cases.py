"""Prompt-assembly cases shared by oracle/make_golden.py (which runs the REFERENCE's function on them) and
tests/test_prompt_cpu.py (which runs the product's).  TEST INFRASTRUCTURE -- never imported by the product."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "faster-qwen3-tts_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def build_base(seed: int = 0):
    """tiny synthetic model, fp32, CPU -- deterministic (seeded CPU generator)"""
    from faster_qwen3_tts import synthetic
    cfg = synthetic.make_config("tiny")
    base = synthetic.build_base_model(cfg, None, seed=seed, dtype=torch.float32, device="cpu")
    base.syn_cfg = cfg
    return base


def cases(base):
    """name -> kwargs of _build_talker_inputs_local / build_talker_inputs (everything except `m`)"""
    tok = lambda s: base._tokenize_texts([base._build_assistant_text(s)])[0]
    ref = lambda s: base._tokenize_texts([base._build_ref_text(s)])[0]
    ins = lambda s: base._tokenize_texts([base._build_instruct_text(s)])[0]
    xv = base._prompt_items_to_voice_clone_prompt(base.create_voice_clone_prompt("a.wav", "", x_vector_only_mode=True))
    icl_long = base._prompt_items_to_voice_clone_prompt(base.create_voice_clone_prompt(([0.0] * 24000 * 3, 24000), "ref"))
    icl_short = base._prompt_items_to_voice_clone_prompt(base.create_voice_clone_prompt(([0.0] * 12000, 24000), "ref"))
    t_short, t_long = "Short parity test.", "The quick brown fox jumps over the lazy dog, twice, and then once more for good measure."
    out = {}
    out["custom_speaker_stream"] = dict(input_ids=[tok(t_short)], ref_ids=[None], voice_clone_prompt=None, languages=["English"],
                                        speakers=["Vivian"], non_streaming_mode=False, instruct_ids=[None])
    out["custom_dialect_auto_nonstream_instruct"] = dict(
        input_ids=[tok(t_long)], ref_ids=[None], voice_clone_prompt=None, languages=["Auto"], speakers=["uncle_fu"],
        non_streaming_mode=True, instruct_ids=[ins("Speak slowly and warmly.")])
    out["custom_nospeaker_auto_stream"] = dict(input_ids=[tok(t_short)], ref_ids=[None], voice_clone_prompt=None,
                                               languages=["auto"], speakers=[None], non_streaming_mode=False, instruct_ids=None)
    out["clone_xvec_stream"] = dict(input_ids=[tok(t_long)], ref_ids=[None], voice_clone_prompt=xv, languages=["German"],
                                    speakers=None, non_streaming_mode=False, instruct_ids=[None])
    out["clone_xvec_nonstream"] = dict(input_ids=[tok(t_short)], ref_ids=[None], voice_clone_prompt=xv, languages=["English"],
                                       speakers=None, non_streaming_mode=True, instruct_ids=[None])
    out["clone_icl_stream_text_shorter"] = dict(input_ids=[tok(t_short)], ref_ids=[ref("A reference line.")],
                                                voice_clone_prompt=icl_long, languages=["English"], speakers=None,
                                                non_streaming_mode=False, instruct_ids=[None])
    out["clone_icl_stream_text_longer"] = dict(input_ids=[tok(t_long)], ref_ids=[ref("A reference line that keeps going.")],
                                               voice_clone_prompt=icl_short, languages=["English"], speakers=None,
                                               non_streaming_mode=False, instruct_ids=[None])
    out["clone_icl_nonstream"] = dict(input_ids=[tok(t_short)], ref_ids=[ref("A reference line.")], voice_clone_prompt=icl_short,
                                      languages=["Japanese"], speakers=None, non_streaming_mode=True,
                                      instruct_ids=[ins("Whisper.")])
    out["custom_batch2_left_pad"] = dict(input_ids=[tok(t_short), tok(t_long)], ref_ids=[None, None], voice_clone_prompt=None,
                                         languages=["English", "Chinese"], speakers=["Ryan", None], non_streaming_mode=False,
                                         instruct_ids=[None, ins("Be brief.")])
    return out
