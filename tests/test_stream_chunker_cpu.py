"""The host-driven (duck-type) path of `fast_generate` / `fast_generate_streaming` -- the loop every non-engine graph
pair goes through -- against the seven runs recorded from the REFERENCE's own schedulers (tests/golden/loop.npz,
oracle/make_golden.py): same doubles (the oracle behind the reference's graph duck types), same noise (torch.multinomial
replaced by the recorded uniforms exactly as the recording did).  Pins codes, chunk boundaries and the `is_final` flags,
including the case the reference flags a FULL chunk final: the cache limit cuts the step after its last frame
(streaming.py:130-132 leaves the loop before the buffer-full check at :158)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import make_golden as MG
from oracle import qwen3_tts_oracle as O


@pytest.mark.parametrize("case", MG.LOOP_CASES, ids=[c[0] for c in MG.LOOP_CASES])
def test_duck_type_streaming_matches_reference_recording(case, golden_dir, monkeypatch):
    from faster_qwen3_tts import sampling
    from faster_qwen3_tts.generate import fast_generate
    from faster_qwen3_tts.streaming import fast_generate_streaming
    gold = np.load(os.path.join(golden_dir, "loop.npz"), allow_pickle=False)
    name, wseed, P, Tt, max_new, min_new, do_sample, pen, max_seq, chunk, boost, nseed = case
    cfg = O.cfg_tiny()
    om = O.OracleModel(cfg, O.make_weights(cfg, seed=wseed, eos_boost=boost))
    tie, tth, tpe = O.make_inputs(cfg, P, Tt, seed=wseed)
    uniforms = np.random.default_rng(nseed).random((max_new + 1, 16), dtype=np.float32)
    sp_pred = O.SamplingParams(do_sample=do_sample, repetition_penalty=1.0)
    conf = types.SimpleNamespace(codec_eos_token_id=cfg.codec_eos_token_id, num_code_groups=cfg.num_code_groups,
                                 vocab_size=cfg.talker.vocab_size)
    proxy = MG._TorchProxy()
    monkeypatch.setattr(sampling, "torch", proxy)
    kw = dict(talker_input_embeds=tie[None], attention_mask=torch.ones(1, P, dtype=torch.long),
              trailing_text_hiddens=tth[None], tts_pad_embed=tpe[None, None], config=conf, max_new_tokens=max_new,
              min_new_tokens=min_new, temperature=0.9, top_k=50, top_p=1.0, do_sample=do_sample, repetition_penalty=pen)
    with torch.inference_mode():
        proxy.uniforms = [float(x) for x in uniforms[:, 0]]
        chunks, finals, totals = [], [], []
        for c, t in fast_generate_streaming(talker=MG._Talker(om), predictor_graph=MG._PredGraph(om, sp_pred, uniforms),
                                            talker_graph=MG._TalkerGraph(om, max_seq), chunk_size=chunk, **kw):
            chunks.append(c)
            finals.append(int(t["is_final"]))
            totals.append(t["total_steps_so_far"])
            assert sorted(t.keys()) == ["chunk_index", "chunk_steps", "decode_ms", "is_final", "prefill_ms", "total_steps_so_far"]
        proxy.uniforms = [float(x) for x in uniforms[:, 0]]
        codes, timing = fast_generate(talker=MG._Talker(om), predictor_graph=MG._PredGraph(om, sp_pred, uniforms),
                                      talker_graph=MG._TalkerGraph(om, max_seq), **kw)
    want = gold[name + "_codes"]
    got = torch.cat(chunks).numpy() if chunks else np.zeros((0, 16), dtype=np.int64)
    assert np.array_equal(got, want)
    assert [c.shape[0] for c in chunks] == gold[name + "_chunks"].tolist()
    assert finals == gold[name + "_final"].tolist()
    assert totals == np.cumsum(gold[name + "_chunks"]).tolist()
    assert np.array_equal(codes.numpy() if codes is not None else np.zeros((0, 16), dtype=np.int64), want)
    assert sorted(timing.keys()) == ["decode_s", "ms_per_step", "prefill_ms", "steps", "steps_per_s"]
