"""bf16 parity of the paths the benchmark actually runs (VERDICT r1 "weak" 1-5): the split-key TMA-staged attention,
the fused loop at the benchmark prompt length, the hand-written prefill at full size and with left padding.

The oracle is hosted on the GPU for these full-size cases (plain torch eager in bf16 = the reference's own arithmetic
on this device); all engine calls go through the C ABI.  bf16 results are compared with stated tolerances; token
streams are compared as first-divergence reports (the reference itself treats bf16 token parity as hardware-fragile,
tests/test_e2e_parity.py:236-313 switches to fp32 for exact tokens -- that bar is held by the fp32 tests)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import qwen3_tts_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from util_models import Pair, report
    from faster_qwen3_tts.engine import SamplingParams
    from faster_qwen3_tts.generate import fast_generate


def _cfg(size):
    return O.cfg_1p7b() if size == "1.7B" else O.cfg_0p6b()


@pytest.fixture(scope="module", params=["0.6B", "1.7B"])
def full_pair(request):
    p = Pair(_cfg(request.param), seed=3, dtype=torch.bfloat16, max_seq_len=2048, oracle_device="cuda")
    # the same (bf16-valued) weights evaluated in fp32: |oracle_bf16 - oracle_fp32| per tensor is the rounding envelope of
    # a VALID bf16 implementation; the engine has to stay within a small multiple of it
    p.om32 = O.OracleModel(p.cfg, {k: v.float().cuda() for k, v in p.W.items()})
    return p


def _rel(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs().max().item()
    mag = ref.abs().max().item()
    return err, mag


def test_split_attention_step_vs_oracle_layerwise(full_pair):
    """talker step at positions >= 192 (keys split over CTAs, K/V slices TMA-staged) against the ORACLE, layer by layer.
    Bar per tensor: engine-vs-bf16-oracle error <= 3 x the bf16 rounding envelope (bf16 oracle vs the same weights in
    fp32) + 1 % of the tensor's range -- i.e. the engine is as close to the bf16 oracle as one valid bf16 evaluation is
    to another; the final hidden state additionally has an absolute bar."""
    p = full_pair
    cfg = p.cfg.talker
    g = torch.Generator().manual_seed(17)
    worst = {}
    for pos in (192, 333, 1100, 2040):
        cache, cache32 = O.KVCache(cfg.num_hidden_layers), O.KVCache(cfg.num_hidden_layers)
        for l in range(cfg.num_hidden_layers):
            k = torch.randn(cfg.num_key_value_heads, pos, 128, generator=g).to(torch.bfloat16).cuda()
            v = (torch.randn(cfg.num_key_value_heads, pos, 128, generator=g) * 0.7).to(torch.bfloat16).cuda()
            cache.k[l], cache.v[l] = k, v
            cache32.k[l], cache32.v[l] = k.float(), v.float()
            p.engine.import_kv(l, k, v)
        p.engine.set_generation_state(0, 0)
        x = (torch.randn(cfg.hidden_size, generator=g) * 0.5).to(torch.bfloat16)
        dbg, dbg32 = {}, {}
        with torch.inference_mode():
            ref = p.om.talker_step(x.cuda(), pos, cache, dbg=dbg)
            ref32 = p.om32.talker_step(x.float().cuda(), pos, cache32, dbg=dbg32)
        p.engine.debug_enable(True)
        got = p.engine.talker_step(x.cuda(), pos)
        torch.cuda.synchronize()
        d = p.engine.debug_layers("t", 1)
        p.engine.debug_enable(False)
        for li in range(cfg.num_hidden_layers):
            for key in ("qkv", "attn", "x1", "act", "x"):
                err, mag = _rel(key, d[f"L{li}.{key}"], dbg[f"L{li}.{key}"])
                env, _ = _rel(key, dbg[f"L{li}.{key}"], dbg32[f"L{li}.{key}"])
                w = worst.setdefault(key, [0.0, 0.0, 0.0])
                if err / (env + 0.01 * mag + 1e-6) > w[0]:
                    worst[key] = [err / (env + 0.01 * mag + 1e-6), err, env]
                assert err <= 3.0 * env + 0.01 * mag + 1e-3, (pos, li, key, err, env, mag)
        err, mag = _rel("hidden", got, ref)
        env, _ = _rel("hidden", ref, ref32)
        print(f"pos {pos}: final hidden engine-vs-bf16-oracle max|d|={err:.4f}  bf16-vs-fp32 oracle envelope {env:.4f}  max|ref|={mag:.3f}")
        assert err <= 3.0 * env + 0.01 * mag and err < 0.25
    print("worst (err / (envelope + 1% range), err, envelope) per tensor kind:", {k: tuple(round(x, 4) for x in v) for k, v in worst.items()})


def test_fused_bf16_loop_at_bench_prompt_vs_oracle(full_pair):
    """P=232 (the benchmark prompt): (a) greedy fused loop against the bf16 oracle, first divergence reported;
    (b) teacher-forced: the oracle's own talker inputs are stepped through the engine (K3 prefill + split attention at
    positions 232..) and hidden / codec_head logits must stay within tolerance for every one of 32 frames."""
    p = full_pair
    cfg = p.cfg
    P, Tt, n = 232, 25, 32
    tie, tth, tpe = O.make_inputs(cfg, P, Tt, seed=9, dtype=torch.bfloat16)
    trace = []
    with torch.inference_mode():
        want = O.generate(p.om, tie.cuda(), tth.cuda(), tpe.cuda(), max_new_tokens=n, min_new_tokens=n,
                          sp_talker=O.SamplingParams(do_sample=False, repetition_penalty=1.05),
                          sp_pred=O.SamplingParams(do_sample=False), max_seq_len=2048, trace=trace)
    p.pg.do_sample = False
    codes, _ = fast_generate(p.talker, tie[None].cuda(), torch.ones(1, P, dtype=torch.long).cuda(), tth[None].cuda(),
                             tpe[None, None].cuda(), p.config, p.pg, p.tg, max_new_tokens=n, min_new_tokens=n,
                             do_sample=False, repetition_penalty=1.05)
    p.pg.do_sample = True
    codes = codes.cpu()
    assert codes.shape == want.shape
    rows_equal = (codes == want).all(dim=1)
    first_div = int((~rows_equal).nonzero()[0]) if (~rows_equal).any() else n
    cb0_equal = (codes[:, 0] == want[:, 0])
    first_cb0 = int((~cb0_equal).nonzero()[0]) if (~cb0_equal).any() else n
    print(f"fused bf16 greedy vs oracle: first frame with any differing code {first_div}/{n}, first differing cb0 {first_cb0}/{n}, "
          f"codes equal {(codes == want).float().mean().item():.3f}")
    # a divergence is legitimate only as a near-tie flip: at the first differing code the oracle's own top-2 logit margin
    # must be within twice the logit tolerance established below (greedy bf16 decoding is chaotic after the first flip;
    # the reference's own parity suite therefore pins tokens in fp32 only, tests/test_e2e_parity.py:236-313)
    if first_div < n:
        col = int((codes[first_div] != want[first_div]).nonzero()[0])
        if col == 0 and first_div == 0:
            margin = 0.0     # the first cb0 comes from the prefill logits (covered by the prefill test)
        elif col == 0:
            lg = trace[first_div - 1]["logits"].clone()
            lg[cfg.talker.vocab_size - 1024:] = float("-inf")
            t2 = torch.topk(lg, 2).values
            margin = float(t2[0] - t2[1])
        else:
            margin = trace[first_div]["pred_margins"][col - 1] if first_div < len(trace) else 0.0
        print(f"first divergence: frame {first_div}, codebook {col}, oracle top-2 margin there {margin:.4f}")
        assert margin < 0.6, (first_div, col, margin)
    # ---- (b) teacher-forced steps on top of the engine's own prefill
    lg, hid = p.engine.prefill(tie.cuda(), 0)
    p.engine.set_generation_state(0, 0)
    head = p.W["talker.codec_head.weight"].cuda()
    worst_h = worst_l = 0.0
    agree = 0
    for s, tr in enumerate(trace):
        got = p.engine.talker_step(tr["x_raw"].cuda(), tr["position"])
        lgt = F.linear(got, head).float().cpu()
        eh = (got.float().cpu() - tr["hidden"]).abs().max().item()
        el = (lgt - tr["logits"]).abs().max().item()
        worst_h, worst_l = max(worst_h, eh), max(worst_l, el)
        agree += int(int(lgt[: cfg.talker.vocab_size - 1024].argmax()) == int(tr["logits"][: cfg.talker.vocab_size - 1024].argmax()))
    print(f"teacher-forced {len(trace)} steps at positions {trace[0]['position']}..{trace[-1]['position']}: "
          f"max|d hidden|={worst_h:.4f} max|d logits|={worst_l:.4f} argmax agreement {agree}/{len(trace)}")
    assert len(trace) >= 31
    assert worst_h < 0.2      # post-norm hidden, |x| ~ 3-4
    assert worst_l < 0.6      # logits, |l| ~ 4-5 (head std 0.08 over H inputs)
    assert agree >= len(trace) - 3


def test_native_prefill_full_size_vs_oracle(full_pair):
    """K3 at the benchmark shape (P=232) and with LEFT PADDING: KV rows, past_hidden and logits against the oracle."""
    p = full_pair
    cfg = p.cfg
    for P, pad in ((232, 0), (96, 13)):
        tie, tth, tpe = O.make_inputs(cfg, P, 3, seed=5 + pad, dtype=torch.bfloat16)
        if pad:
            tie[:pad] = 0
        with torch.inference_mode():
            logits, ph, cache = p.om.talker_prefill(tie.cuda(), n_left_pad=pad)
        lg, hid = p.engine.prefill(tie.cuda(), pad)
        eh, mh = _rel("hidden", hid, ph)
        el, ml = _rel("logits", lg, logits)
        print(f"P={P} pad={pad}: past_hidden max|d|={eh:.4f} (|ref| {mh:.2f})  logits max|d|={el:.4f} (|ref| {ml:.2f})")
        assert eh < 0.2 and el < 0.6
        assert int(lg.float()[: cfg.talker.vocab_size - 1024].argmax()) == int(logits.float()[: cfg.talker.vocab_size - 1024].argmax()) or el < 0.3
        for l in (0, cfg.talker.num_hidden_layers // 2, cfg.talker.num_hidden_layers - 1):
            k, v = p.engine.export_kv(l, P)
            ek, mk = _rel("k", k[:, pad:], cache.k[l][:, pad:])
            ev, mv = _rel("v", v[:, pad:], cache.v[l][:, pad:])
            print(f"   layer {l}: K max|d|={ek:.4f} (|ref| {mk:.2f})  V max|d|={ev:.4f} (|ref| {mv:.2f})")
            assert ek <= 0.04 * mk + 2e-2 and ev <= 0.04 * mv + 2e-2
        # a decode step on the natively written cache (left pad + rope delta on device)
        x = (torch.randn(cfg.talker.hidden_size, generator=torch.Generator().manual_seed(8)) * 0.5).to(torch.bfloat16)
        p.engine.set_generation_state(pad, -pad)
        got = p.engine.talker_step(x.cuda(), P)
        with torch.inference_mode():
            ref = p.om.talker_step(x.cuda(), P, cache, n_left_pad=pad, rope_delta=-pad)
        es, ms = _rel("step", got, ref)
        print(f"   decode step after prefill: max|d|={es:.4f} (|ref| {ms:.2f})")
        assert es < 0.2
    p.engine.set_generation_state(0, 0)


def test_predictor_bf16_pass0_logits_and_codes(full_pair):
    """bf16 predictor frame: pass-0 logits (from the dumped last-layer residual stream) within tolerance of the oracle,
    and the greedy code of every pass equal wherever the oracle's top-2 margin exceeds twice that tolerance."""
    p = full_pair
    cfg = p.cfg
    g = torch.Generator().manual_seed(31)
    ph = (torch.randn(cfg.talker.hidden_size, generator=g) * 0.8).to(torch.bfloat16)
    emb = p.W["talker.model.codec_embedding.weight"][41]
    dbg = {}
    with torch.inference_mode():
        ref = p.om.predictor_frame(ph.cuda(), emb.cuda(), O.SamplingParams(do_sample=False), dbg=dbg)
    p.engine.debug_enable(True)
    got = p.engine.predictor_run(torch.stack((ph, emb)).cuda(), SamplingParams(do_sample=False))
    torch.cuda.synchronize()
    d = p.engine.debug_layers("p", 2)
    p.engine.debug_enable(False)
    Lp = cfg.predictor.num_hidden_layers
    x_last = d[f"L{Lp - 1}.x"][1].to(torch.bfloat16)
    w = p.W["talker.code_predictor.model.norm.weight"]
    hidn = O.rms_norm(x_last, w, cfg.predictor.rms_norm_eps)
    lg = F.linear(hidn, p.W["talker.code_predictor.lm_head.0.weight"]).float()
    ref_lg = dbg["pred.logits0"].cpu()
    err = (lg - ref_lg).abs().max().item()
    top2 = torch.topk(ref_lg, 2).values
    print(f"predictor pass-0 logits max|d|={err:.4f}, oracle top-2 margin {float(top2[0] - top2[1]):.4f}; codes got {got.tolist()} ref {ref}")
    assert err < 0.25
    if float(top2[0] - top2[1]) > 2 * 0.25:
        assert int(got[0]) == ref[0]
    assert sum(int(a == b) for a, b in zip(got.tolist(), ref)) >= 8   # later passes depend on earlier codes


def test_full_size_fp32_gpu_hosted_oracle_tokens_bit_exact():
    """0.6B geometry, fp32, 40 frames greedy with left padding 7 / rope delta -7 on device: exact tokens
    (the reference's exact-parity protocol, tests/test_e2e_parity.py:236-313)."""
    cfg = O.cfg_0p6b()
    p = Pair(cfg, seed=4, dtype=torch.float32, max_seq_len=256, oracle_device="cuda")
    p.pg.do_sample = False
    P, pad, n = 47, 7, 40
    tie, tth, tpe = O.make_inputs(cfg, P, 4, seed=2)
    tie[:pad] = 0
    tam = torch.ones(1, P, dtype=torch.long)
    tam[0, :pad] = 0
    with torch.inference_mode():
        want = O.generate(p.om, tie.cuda(), tth.cuda(), tpe.cuda(), max_new_tokens=n,
                          sp_talker=O.SamplingParams(do_sample=False, repetition_penalty=1.05),
                          sp_pred=O.SamplingParams(do_sample=False), max_seq_len=256, n_left_pad=pad)
    codes, _ = fast_generate(p.talker, tie[None].cuda(), tam.cuda(), tth[None].cuda(), tpe[None, None].cuda(), p.config,
                             p.pg, p.tg, max_new_tokens=n, min_new_tokens=2, do_sample=False, repetition_penalty=1.05)
    codes = codes.cpu()
    print("rows equal:", int((codes == want[: codes.shape[0]]).all(dim=1).sum()), "of", want.shape[0])
    assert torch.equal(codes, want)
