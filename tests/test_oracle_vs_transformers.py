"""The oracle's layer arithmetic against the Hugging Face eager Qwen3 decoder (`transformers.models.qwen3`).

The talker and the code predictor of Qwen3-TTS are Qwen3 decoder stacks (RMSNorm, q/k-norm, RoPE theta 1e6, GQA,
SwiGLU) whose real implementation lives in un-vendored `qwen-tts` / `transformers`.
`transformers` IS in the image, so the oracle's `run_stack` -- the thing every GPU parity test is measured against --
is pinned here to an independent third-party implementation of the same block: same weights, prefill over a prompt
(causal mask), then cached single-token steps (the decode path), in fp32 and in bf16 (which checks the rounding points
the oracle and the device kernel emulate).  The mRoPE of the talker uses three identical position streams, which is
plain RoPE (talker_graph.py:53,210-211).
"""
import pytest
import torch

from oracle import qwen3_tts_oracle as O

tf = pytest.importorskip("transformers")


def _hf_stack(sc: O.StackCfg, W, prefix, dtype):
    from transformers.models.qwen3 import Qwen3Config, Qwen3Model
    cfg = Qwen3Config(vocab_size=32, hidden_size=sc.hidden_size, intermediate_size=sc.intermediate_size,
                      num_hidden_layers=sc.num_hidden_layers, num_attention_heads=sc.num_attention_heads,
                      num_key_value_heads=sc.num_key_value_heads, head_dim=sc.head_dim, rms_norm_eps=sc.rms_norm_eps,
                      rope_theta=sc.rope_theta, max_position_embeddings=512, attention_bias=False,
                      tie_word_embeddings=False, use_sliding_window=False)
    cfg._attn_implementation = "eager"
    m = Qwen3Model(cfg).eval()
    sd = {k[len(prefix) + 1:]: v for k, v in W.items() if k.startswith(prefix + ".layers.") or k == prefix + ".norm.weight"}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and missing == ["embed_tokens.weight"], (missing, unexpected)
    return m.to(dtype)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 6e-2)])
@pytest.mark.parametrize("which", ["talker", "predictor"])
def test_run_stack_matches_hf_qwen3_prefill_and_cached_steps(which, dtype, tol):
    cfg = O.cfg_tiny()
    sc = cfg.talker if which == "talker" else cfg.predictor
    prefix = "talker.model" if which == "talker" else "talker.code_predictor.model"
    W = O.make_weights(cfg, seed=3, dtype=dtype)
    hf = _hf_stack(sc, W, prefix, dtype)
    rope = O.rope_tables(sc.head_dim, sc.rope_theta, 512)
    g = torch.Generator().manual_seed(5)
    P, steps = 9, 3
    x = (torch.randn(P + steps, sc.hidden_size, generator=g) * 0.5).to(dtype)
    with torch.inference_mode():
        cache = O.KVCache(sc.num_hidden_layers)
        mine = [O.run_stack(W, prefix, sc, x[:P], torch.arange(P), cache, rope)]
        out = hf(inputs_embeds=x[None, :P], use_cache=True)
        theirs = [out.last_hidden_state[0]]
        kv = out.past_key_values
        for s in range(steps):
            mine.append(O.run_stack(W, prefix, sc, x[P + s:P + s + 1], torch.tensor([P + s]), cache, rope))
            out = hf(inputs_embeds=x[None, P + s:P + s + 1], past_key_values=kv, use_cache=True)
            kv = out.past_key_values
            theirs.append(out.last_hidden_state[0])
    a, b = torch.cat(mine).float(), torch.cat(theirs).float()
    err = (a - b).abs().max().item()
    print(f"{which} {dtype}: max abs diff {err:.3e} over {a.shape[0]} positions (|x| ~ {b.abs().mean():.2f})")
    assert a.shape == b.shape and err < tol


def test_kv_cache_rows_match_hf_cache():
    """the K rows the oracle appends (post k-norm, post RoPE) are what HF caches -- this is the layout the engine's
    fq3_import_kv expects (talker_graph.py:153-170)"""
    cfg = O.cfg_tiny()
    sc, prefix = cfg.talker, "talker.model"
    W = O.make_weights(cfg, seed=4, dtype=torch.float32)
    hf = _hf_stack(sc, W, prefix, torch.float32)
    x = torch.randn(6, sc.hidden_size, generator=torch.Generator().manual_seed(1)) * 0.5
    with torch.inference_mode():
        cache = O.KVCache(sc.num_hidden_layers)
        O.run_stack(W, prefix, sc, x, torch.arange(6), cache, O.rope_tables(sc.head_dim, sc.rope_theta, 64))
        kv = hf(inputs_embeds=x[None], use_cache=True).past_key_values
    for li in range(sc.num_hidden_layers):
        layer = kv.layers[li] if hasattr(kv, "layers") else None
        hk, hv = (layer.keys, layer.values) if layer is not None else kv[li]
        ok = cache.k[li].reshape(hk.shape[1], -1, sc.head_dim) if cache.k[li].dim() != 3 else cache.k[li]
        ov = cache.v[li].reshape(hv.shape[1], -1, sc.head_dim) if cache.v[li].dim() != 3 else cache.v[li]
        assert (ok - hk[0]).abs().max() < 1e-5 and (ov - hv[0]).abs().max() < 1e-5


def test_predictor_15_step_loop_matches_hf_code_predictor():
    """The code predictor's per-frame loop (predictor_graph.py:115-167: 2-token prefill [past_hidden, embed(cb0)], then
    14 single-token steps where step i embeds the previous code with table i-1 and reads head i) against the Hugging Face
    Qwen3-Omni talker code predictor driven step by step with the same weights, greedy, fp32: identical codes and
    logits.  (That analogue has no small_to_mtp_projection, so the geometry used here has equal talker / predictor
    widths -- the 0.6B case, where the projection is the identity.)"""
    import dataclasses
    from transformers.models.qwen3_omni_moe import configuration_qwen3_omni_moe as Cf, modeling_qwen3_omni_moe as M
    base = O.cfg_tiny()
    cfg = dataclasses.replace(base, talker=dataclasses.replace(base.talker, hidden_size=base.predictor.hidden_size),
                              has_mtp_projection=False)
    pc = cfg.predictor
    W = O.make_weights(cfg, seed=6, dtype=torch.float32)
    hcfg = Cf.Qwen3OmniMoeTalkerCodePredictorConfig(
        vocab_size=pc.vocab_size, hidden_size=pc.hidden_size, intermediate_size=pc.intermediate_size,
        num_hidden_layers=pc.num_hidden_layers, num_attention_heads=pc.num_attention_heads,
        num_key_value_heads=pc.num_key_value_heads, head_dim=pc.head_dim, rms_norm_eps=pc.rms_norm_eps,
        rope_parameters={"rope_theta": pc.rope_theta, "rope_type": "default"}, max_position_embeddings=64,
        num_code_groups=cfg.num_code_groups)
    hcfg._attn_implementation = "eager"
    hf = M.Qwen3OmniMoeTalkerCodePredictorModelForConditionalGeneration(hcfg).eval()
    pre = "talker.code_predictor."
    missing, unexpected = hf.load_state_dict({k[len(pre):]: v for k, v in W.items() if k.startswith(pre)}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    om = O.OracleModel(cfg, W)
    g = torch.Generator().manual_seed(2)
    for trial in range(3):
        past_hidden = torch.randn(pc.hidden_size, generator=g)
        cb0 = int(torch.randint(0, pc.vocab_size, (1,), generator=g))
        last_id_hidden = om.codec_embed(cb0)
        dbg = {}
        with torch.inference_mode():
            mine = om.predictor_frame(past_hidden, last_id_hidden, O.SamplingParams(do_sample=False))
            out = hf(inputs_embeds=torch.stack([past_hidden, last_id_hidden])[None], use_cache=True)
            theirs = [int(out.logits[0, -1].argmax())]
            for i in range(1, cfg.num_code_groups - 1):
                out = hf(input_ids=torch.tensor([[theirs[-1]]]), past_key_values=out.past_key_values, use_cache=True,
                         generation_steps=i)
                theirs.append(int(out.logits[0, -1].argmax()))
        assert mine == theirs, (trial, mine, theirs)


def test_three_identical_mrope_streams_equal_plain_rope_tables():
    """talker_graph.py:53,210-211 feeds the same position on the three mRoPE axes; with the interleaved-mRoPE module
    of the Hugging Face talker that must equal the plain RoPE tables the oracle (`rope_tables`) and the engine's
    host side (`weights.rope_tables`) build -- including a rope_delta offset (negative positions are clamped by the
    callers, not here)."""
    from transformers.models.qwen3_omni_moe import configuration_qwen3_omni_moe as Cf, modeling_qwen3_omni_moe as M
    from oracle import prompt_cases  # noqa: F401  (puts the product package on sys.path)
    from faster_qwen3_tts.weights import rope_tables as product_tables
    tcfg = Cf.Qwen3OmniMoeTalkerTextConfig(hidden_size=512, num_attention_heads=4, num_key_value_heads=2, head_dim=128,
                                           rope_parameters={"rope_theta": 1_000_000.0, "rope_type": "default",
                                                            "mrope_section": [24, 20, 20]})
    rot = M.Qwen3OmniMoeTalkerRotaryEmbedding(tcfg)
    pos = torch.tensor([[0, 1, 2, 17, 255, 1000, 2047]]) + 3          # "+3": a rope_delta
    cos, sin = rot(torch.zeros(1, 1, 128), pos[None].expand(3, -1, -1))
    oc, os_ = O.rope_tables(128, 1_000_000.0, 4096)
    assert torch.equal(cos[0], oc[pos[0]]) and torch.equal(sin[0], os_[pos[0]])
    pc, ps = product_tables(1_000_000.0, 4096, 128)
    assert torch.equal(torch.as_tensor(pc)[pos[0]].float(), oc[pos[0]]) and torch.equal(torch.as_tensor(ps)[pos[0]].float(), os_[pos[0]])
