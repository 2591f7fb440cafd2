"""GPU parity of the BATCHED decode path (BASELINE config 4; SURVEY section 8(e): "up to 32 rows share one weight
stream"): B request slots advance in one persistent-kernel launch per chunk.

Bars:
  * fp32: every row's codes bit-exact against B INDEPENDENT oracle runs, with mixed prompt lengths, LEFT PADDING and the
    matching non-zero rope delta (the reference's batch layout, model.py:774-787 / talker_graph.py:177-187), sampled
    and greedy, rows finishing at different times (EOS / max_new_tokens);
  * bf16 (the dtype of every reported number): every row of a batched run bit-identical to the same request run alone
    through the single-sequence kernel (tiny and 1.7B geometry), so the bf16 parity established for the single-sequence
    path carries over to the batched path row by row.
All calls go through the C ABI (fq3_begin_request(slot) / fq3_decode_chunk(slots[], n_slots, ...))."""
import numpy as np
import pytest
import torch

from oracle import qwen3_tts_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from util_models import Pair
    from faster_qwen3_tts.batching import BatchScheduler, fast_generate_batch, fast_generate_streaming_batch
    from faster_qwen3_tts.generate import fast_generate


def _left_padded_batch(cfg, lens, Tt_list, seed, dtype):
    """Rows with different prompt lengths, left-padded with zero rows like the reference's batch builder."""
    Pmax, Tmax = max(lens), max(max(Tt_list), 1)
    H = cfg.talker.hidden_size
    tie = torch.zeros(len(lens), Pmax, H, dtype=dtype)
    tam = torch.zeros(len(lens), Pmax, dtype=torch.long)
    tth = torch.zeros(len(lens), Tmax, H, dtype=dtype)
    rows = []
    tpe = None
    for b, (P, Tt) in enumerate(zip(lens, Tt_list)):
        e, t, pad = O.make_inputs(cfg, P, Tt, seed=seed + b, dtype=dtype)
        if tpe is None:
            tpe = pad
        tie[b, Pmax - P:] = e
        tam[b, Pmax - P:] = 1
        tth[b] = tpe            # rows beyond a request's own trailing text carry tts_pad_embed (model.py:789-803)
        if Tt:
            tth[b, :Tt] = t
        rows.append((e, t))
    return tie, tam, tth, tpe, rows


def _oracle_rows(p, tie, tam, tth, tpe, uniforms, sp_t, sp_p, max_new, min_new, max_seq_len):
    want = []
    # tiny tensors: intra-op threading only adds synchronisation (a 32-row oracle run took minutes on a many-core box)
    nthr = torch.get_num_threads()
    torch.set_num_threads(1)
    for b in range(tie.shape[0]):
        pad = int((tam[b] == 0).sum())
        with torch.inference_mode():
            want.append(O.generate(p.om, tie[b], tth[b], tpe, max_new_tokens=max_new[b] if isinstance(max_new, list) else max_new,
                                   min_new_tokens=min_new, sp_talker=sp_t, sp_pred=sp_p, max_seq_len=max_seq_len,
                                   uniforms=None if uniforms is None else uniforms[b], n_left_pad=pad))
    torch.set_num_threads(nthr)
    return want


@pytest.mark.parametrize("B", [2, 8, 32])
@pytest.mark.parametrize("do_sample", [False, True])
def test_batched_fp32_rows_match_independent_oracle_runs(B, do_sample):
    cfg = O.cfg_tiny()
    p = Pair(cfg, seed=B, dtype=torch.float32, max_seq_len=96, eos_boost=3.0 if do_sample else 1.0, max_batch=B)
    p.pg.do_sample = do_sample
    rng = np.random.default_rng(100 + B)
    lens = [int(x) for x in rng.integers(5, 30, size=B)]
    lens[0] = max(lens) + 3           # row 0 unpadded, all others left-padded by different amounts
    Tts = [int(x) for x in rng.integers(0, 6, size=B)]
    tie, tam, tth, tpe, _ = _left_padded_batch(cfg, lens, Tts, seed=7 * B, dtype=torch.float32)
    assert int((tam == 0).sum()) > 0
    n = 14
    uniforms = rng.random((B, n + 1, 16), dtype=np.float32) if do_sample else None
    sp_t = O.SamplingParams(do_sample=do_sample, repetition_penalty=1.05)
    sp_p = O.SamplingParams(do_sample=do_sample)
    want = _oracle_rows(p, tie, tam, tth, tpe, uniforms, sp_t, sp_p, n, 2, 96)
    got, timing = fast_generate_batch(
        p.talker, tie.cuda(), tam.cuda(), tth.cuda(), tpe[None, None].cuda(), p.config, p.pg, p.tg, max_new_tokens=n,
        min_new_tokens=2, do_sample=do_sample, repetition_penalty=1.05,
        uniforms=None if uniforms is None else torch.from_numpy(uniforms).cuda(), launch_frames=5)
    lens_got = [0 if g is None else g.shape[0] for g in got]
    print("B", B, "frames per row:", lens_got, "oracle:", [w.shape[0] for w in want])
    bad = [b for b in range(B) if lens_got[b] != want[b].shape[0] or (lens_got[b] and not torch.equal(got[b].cpu(), want[b]))]
    assert not bad, bad
    assert set(timing) == {"prefill_ms", "decode_s", "steps", "ms_per_step", "steps_per_s"}
    if do_sample:
        assert len(set(lens_got)) > 1 or B == 2   # eos_boost: rows stop at different frames


def test_batched_fp32_max_seq_len_and_max_new_stops():
    """rows that hit the cache limit (generate.py:175-177: frame emitted, no further talker step), the max_new_tokens
    bound and a long run, side by side in one batch (different prompt lengths, submitted through the scheduler)"""
    cfg = O.cfg_tiny()
    S = 40
    p = Pair(cfg, seed=3, dtype=torch.float32, max_seq_len=S, max_batch=4)
    p.pg.do_sample = False
    specs = [(34, 2, 20), (12, 0, 4), (10, 3, 20), (30, 1, 20)]   # (P, Tt, max_new)
    sp_t = O.SamplingParams(do_sample=False, repetition_penalty=1.05)
    sp_p = O.SamplingParams(do_sample=False)
    sched = BatchScheduler(p.engine, p.talker, p.config, p.pg, p.tg)
    want, got = [], {}
    for i, (P, Tt, n) in enumerate(specs):
        e, t, pad = O.make_inputs(cfg, P, Tt, seed=60 + i)
        with torch.inference_mode():
            want.append(O.generate(p.om, e, t, pad, max_new_tokens=n, min_new_tokens=2, sp_talker=sp_t, sp_pred=sp_p,
                                   max_seq_len=S))
        sched.submit(e[None].cuda(), torch.ones(1, P, dtype=torch.long).cuda(), t[None].cuda(), pad[None, None].cuda(),
                     tag=i, max_new_tokens=n, min_new_tokens=2, do_sample=False, repetition_penalty=1.05)
        got[i] = []
    fins = {}
    while len(sched):
        for rq, codes in sched.step(3):
            got[rq.tag].append(codes.cpu())
            fins[rq.tag] = rq.finished
    for i in range(len(specs)):
        assert torch.equal(torch.cat(got[i]), want[i]), i
    assert want[0].shape[0] == S - 1 - 34 + 1 and fins[0] == 3   # cache limit
    assert want[1].shape[0] == 4 and fins[1] == 1                # max_new_tokens


def _single_runs(p, tie, tam, tth, tpe, uniforms, **kw):
    out = []
    for b in range(tie.shape[0]):
        codes, _ = fast_generate(p.talker, tie[b:b + 1].cuda(), tam[b:b + 1].cuda(), tth[b:b + 1].cuda(),
                                 tpe[None, None].cuda(), p.config, p.pg, p.tg,
                                 uniforms=None if uniforms is None else torch.from_numpy(uniforms[b]).cuda(), **kw)
        out.append(codes.cpu() if codes is not None else None)
    return out


@pytest.mark.parametrize("B", [3, 8, 20])
def test_batched_bf16_rows_bit_identical_to_single_sequence_kernel(B):
    cfg = O.cfg_tiny()
    p = Pair(cfg, seed=1, dtype=torch.bfloat16, max_seq_len=160, eos_boost=2.0, max_batch=B)
    rng = np.random.default_rng(B)
    lens = [int(x) for x in rng.integers(6, 60, size=B)]
    Tts = [int(x) for x in rng.integers(0, 8, size=B)]
    tie, tam, tth, tpe, _ = _left_padded_batch(cfg, lens, Tts, seed=5 * B, dtype=torch.bfloat16)
    n = 24
    uniforms = rng.random((B, n + 1, 16), dtype=np.float32)
    kw = dict(max_new_tokens=n, min_new_tokens=2, do_sample=True, repetition_penalty=1.05)
    want = _single_runs(p, tie, tam, tth, tpe, uniforms, **kw)
    got, _ = fast_generate_batch(p.talker, tie.cuda(), tam.cuda(), tth.cuda(), tpe[None, None].cuda(), p.config, p.pg,
                                 p.tg, uniforms=torch.from_numpy(uniforms).cuda(), launch_frames=8, **kw)
    for b in range(B):
        a, w = got[b], want[b]
        assert (a is None) == (w is None), b
        if a is not None:
            assert a.shape == w.shape and torch.equal(a.cpu(), w), (b, a.shape, w.shape)
    # streaming driver: same rows, reference chunk bookkeeping per row
    parts = [[] for _ in range(B)]
    for items in fast_generate_streaming_batch(p.talker, tie.cuda(), tam.cuda(), tth.cuda(), tpe[None, None].cuda(),
                                               p.config, p.pg, p.tg, chunk_size=8,
                                               uniforms=torch.from_numpy(uniforms).cuda(), **kw):
        for b, codes, tm in items:
            assert set(tm) >= {"chunk_index", "chunk_steps", "prefill_ms", "decode_ms", "total_steps_so_far", "is_final"}
            assert codes.shape[0] == tm["chunk_steps"] <= 8
            parts[b].append(codes.cpu())
    for b in range(B):
        if want[b] is not None:
            assert torch.equal(torch.cat(parts[b]), want[b]), b


def test_batched_bf16_full_size_rows_match_single_sequence_kernel():
    """1.7B geometry (the benchmark model), bf16, 8 rows with different prompt lengths and pads: batched rows ==
    single-sequence runs, frame for frame."""
    cfg = O.cfg_1p7b()
    B = 8
    p = Pair(cfg, seed=2, dtype=torch.bfloat16, max_seq_len=256, max_batch=B)
    lens = [40, 17, 33, 40, 25, 9, 38, 21]
    tie, tam, tth, tpe, _ = _left_padded_batch(cfg, lens, [3, 0, 5, 2, 0, 1, 4, 2], seed=21, dtype=torch.bfloat16)
    n = 10
    uniforms = np.random.default_rng(4).random((B, n + 1, 16), dtype=np.float32)
    kw = dict(max_new_tokens=n, min_new_tokens=n, do_sample=True, repetition_penalty=1.05)
    want = _single_runs(p, tie, tam, tth, tpe, uniforms, **kw)
    got, _ = fast_generate_batch(p.talker, tie.cuda(), tam.cuda(), tth.cuda(), tpe[None, None].cuda(), p.config, p.pg,
                                 p.tg, uniforms=torch.from_numpy(uniforms).cuda(), launch_frames=4, **kw)
    same = [bool(torch.equal(got[b].cpu(), want[b])) for b in range(B)]
    print("rows identical:", same)
    assert all(same)


def test_continuous_batching_join_and_leave_between_chunks():
    """serving pattern: requests join while others are mid-stream and slots are re-used; every
    request's codes equal its stand-alone run."""
    cfg = O.cfg_tiny()
    p = Pair(cfg, seed=6, dtype=torch.bfloat16, max_seq_len=128, max_batch=3)
    reqs = []
    for i, (P, Tt, n) in enumerate([(12, 2, 20), (30, 0, 9), (8, 4, 14), (21, 1, 11), (16, 3, 17)]):
        e, t, pad = O.make_inputs(cfg, P, Tt, seed=40 + i, dtype=torch.bfloat16)
        u = np.random.default_rng(i).random((n + 1, 16), dtype=np.float32)
        reqs.append((e, t, pad, n, u))
    kw = dict(min_new_tokens=2, do_sample=True, repetition_penalty=1.05)
    want = []
    for e, t, pad, n, u in reqs:
        codes, _ = fast_generate(p.talker, e[None].cuda(), torch.ones(1, e.shape[0], dtype=torch.long).cuda(),
                                 t[None].cuda() if t.shape[0] else torch.zeros(1, 0, e.shape[1], dtype=e.dtype).cuda(),
                                 pad[None, None].cuda(), p.config, p.pg, p.tg, max_new_tokens=n,
                                 uniforms=torch.from_numpy(u).cuda(), **kw)
        want.append(codes.cpu())
    sched = BatchScheduler(p.engine, p.talker, p.config, p.pg, p.tg)
    pending = list(range(len(reqs)))
    got = {i: [] for i in pending}
    while pending or len(sched):
        while pending and sched.has_capacity():
            i = pending.pop(0)
            e, t, pad, n, u = reqs[i]
            sched.submit(e[None].cuda(), torch.ones(1, e.shape[0], dtype=torch.long).cuda(),
                         t[None].cuda() if t.shape[0] else torch.zeros(1, 0, e.shape[1], dtype=e.dtype).cuda(),
                         pad[None, None].cuda(), tag=i, max_new_tokens=n, uniforms=torch.from_numpy(u).cuda(), **kw)
        for rq, codes in sched.step(4):
            got[rq.tag].append(codes.cpu())
    for i in range(len(reqs)):
        assert torch.equal(torch.cat(got[i]), want[i]), i


def _gemv_ref(W, x, which):
    """torch reference of one segment GEMV with the engine's rounding points (fp32 accumulation of bf16 products)."""
    xf = x.float()
    if which == 2:
        wg, wu = W
        g = (xf @ wg.float().t()).to(torch.bfloat16).float()
        u = (xf @ wu.float().t()).to(torch.bfloat16).float()
        s = (g / (1.0 + torch.exp(-g))).to(torch.bfloat16).float()
        return (s * u).to(torch.bfloat16).float()
    return (xf @ W.float().t()).to(torch.bfloat16).float()


@pytest.mark.parametrize("size", ["tiny", "0.6B"])
def test_batched_gemv_vs_torch_fp32_reference(size):
    """The kernel every batched pass is built from, alone: y[col][row] = W[row,:] . x[col,:] for every segment kind
    (q|k|v, o_proj, gate/up + SwiGLU, down, heads; FULL / HALF / GU tape tiles) and for column counts that exercise 1-4
    n-groups, ragged groups and the replay for more than 32 columns -- against torch fp32 matmuls of the same bf16
    operands."""
    cfg = O.cfg_tiny() if size == "tiny" else O.cfg_0p6b()
    p = Pair(cfg, seed=5, dtype=torch.bfloat16, max_seq_len=64, max_batch=2)
    g = torch.Generator().manual_seed(3)
    worst = 0.0
    for stack, pre, sc in ((0, "talker.model", cfg.talker), (1, "talker.code_predictor.model", cfg.predictor)):
        layers = sorted({0, sc.num_hidden_layers - 1})
        for l in layers:
            lp = f"{pre}.layers.{l}."
            mats = {0: torch.cat([p.W[lp + f"self_attn.{n}_proj.weight"] for n in "qkv"]), 1: p.W[lp + "self_attn.o_proj.weight"],
                    2: (p.W[lp + "mlp.gate_proj.weight"], p.W[lp + "mlp.up_proj.weight"]), 3: p.W[lp + "mlp.down_proj.weight"]}
            for which, W in mats.items():
                K = (W[0] if which == 2 else W).shape[1]
                for ncols in (1, 3, 8, 13, 20, 32, 40, 64):
                    x = (torch.randn(ncols, K, generator=g) * 0.5).to(torch.bfloat16)
                    got = p.engine.debug_gemv(stack, l, which, x.cuda()).float().cpu()
                    ref = _gemv_ref(tuple(w.cuda() for w in W) if which == 2 else W.cuda(), x.cuda(), which).cpu()
                    err = (got - ref).abs().max().item()
                    mag = ref.abs().max().item()
                    worst = max(worst, err / (mag + 1e-6))
                    # both sides round the fp32 sum to bf16: one bf16 ulp of the largest element + summation-order noise
                    assert err <= 0.012 * mag + 1e-4, (stack, l, which, ncols, err, mag)
        hw = p.W["talker.codec_head.weight"] if stack == 0 else p.W["talker.code_predictor.lm_head.3.weight"]
        x = (torch.randn(17, hw.shape[1], generator=g) * 0.5).to(torch.bfloat16)
        got = p.engine.debug_gemv(stack, 3 if stack else 0, 4, x.cuda()).float().cpu()
        ref = _gemv_ref(hw.cuda(), x.cuda(), 4).cpu()
        assert (got - ref).abs().max().item() <= 0.012 * ref.abs().max().item() + 1e-4
    print(f"{size}: worst relative GEMV error {worst:.2e}")
