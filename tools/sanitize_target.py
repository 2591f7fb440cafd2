"""Workload for compute-sanitizer (tools/sanitize_box.sh): every hand-written kernel family once, on the tiny geometry
so that the instrumented run finishes in minutes -- fused single-sequence decode (fp32 and bf16, greedy and sampled),
split-key talker attention (bf16, >= 192 cached keys), batched decode (3 slots, left padding), K3 prefill and the
codec (front end + waveform stack).  Prints OK lines; any sanitizer finding shows up in the tool's own report."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "faster-qwen3-tts_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import qwen3_tts_oracle as O  # noqa: E402
from util_models import Pair  # noqa: E402
from faster_qwen3_tts.generate import fast_generate  # noqa: E402
from faster_qwen3_tts.batching import fast_generate_batch  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cfg = O.cfg_tiny()


def fused(dtype, do_sample, P, frames, max_seq_len):
    p = Pair(cfg, seed=1, dtype=dtype, max_seq_len=max_seq_len)
    p.pg.do_sample = do_sample
    tie, tth, tpe = O.make_inputs(cfg, P, 2, seed=0, dtype=dtype)
    u = torch.from_numpy(np.random.default_rng(0).random((frames + 1, 16), dtype=np.float32)).cuda() if do_sample else None
    codes, _ = fast_generate(p.talker, tie[None].cuda(), torch.ones(1, P, dtype=torch.long).cuda(), tth[None].cuda(),
                             tpe[None, None].cuda(), p.config, p.pg, p.tg, max_new_tokens=frames, min_new_tokens=frames,
                             do_sample=do_sample, uniforms=u)
    torch.cuda.synchronize()
    print("OK fused", dtype, "sample" if do_sample else "greedy", "P", P, tuple(codes.shape), flush=True)


if which in ("all", "fused"):
    fused(torch.float32, False, 8, 3, 64)
    fused(torch.bfloat16, True, 8, 3, 64)
if which in ("all", "split"):
    fused(torch.bfloat16, True, 200, 2, 256)      # >= 192 cached keys: split-key attention on TMA-staged K/V tiles
if which in ("all", "batch"):
    B = 3
    p = Pair(cfg, seed=2, dtype=torch.bfloat16, max_seq_len=64, max_batch=B)
    H = cfg.talker.hidden_size
    lens = [12, 7, 9]
    tie = torch.zeros(B, 12, H, dtype=torch.bfloat16)
    tam = torch.zeros(B, 12, dtype=torch.long)
    for b, L in enumerate(lens):
        e, t, pad = O.make_inputs(cfg, L, 1, seed=b, dtype=torch.bfloat16)
        tie[b, 12 - L:] = e
        tam[b, 12 - L:] = 1
    tth = pad[None, None].expand(B, 1, H).contiguous()
    got, _ = fast_generate_batch(p.talker, tie.cuda(), tam.cuda(), tth.cuda(), pad[None, None].cuda(), p.config, p.pg, p.tg,
                                 max_new_tokens=3, min_new_tokens=3, do_sample=True, launch_frames=3)
    torch.cuda.synchronize()
    print("OK batch", [tuple(g.shape) for g in got], flush=True)
if which in ("all", "codec"):
    from faster_qwen3_tts.codec import Code2WavConfig, build_codec
    cc = Code2WavConfig(codebook_size=256, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                        intermediate_size=512, sliding_window=8, decoder_dim=512)
    st = build_codec(cc, seed=1, dtype=torch.bfloat16, device="cuda")
    codes = torch.randint(0, 256, (2, 11, 16)).cuda()
    pcm, sr = st.decode({"audio_codes": codes})
    torch.cuda.synchronize()
    print("OK codec", tuple(pcm[0].shape), sr, flush=True)
    # stateful streams: history rows in front of every causal layer, two streams at different positions in one call
    s1, s2 = st.open_stream(), st.open_stream()
    s1.warm(codes[0, :5])
    a = s1.push(codes[0, 5:])
    b = st.push_streams([s1, s2], torch.stack([codes[1, :3], codes[1, :3]]))
    torch.cuda.synchronize()
    whole, _ = st.decode({"audio_codes": codes[:1]})
    print("OK codec streams", tuple(a.shape), [tuple(x.shape) for x in b], "tail equals one-shot:",
          bool(torch.equal(a, whole[0][5 * 1920:])), flush=True)
if which in ("all", "prefill"):
    p = Pair(cfg, seed=3, dtype=torch.bfloat16, max_seq_len=64)
    if p.engine.has_prefill:
        tie, _, _ = O.make_inputs(cfg, 20, 1, seed=0, dtype=torch.bfloat16)
        lg, hid = p.engine.prefill(tie.cuda(), n_left_pad=3)
        torch.cuda.synchronize()
        print("OK prefill", tuple(lg.shape), flush=True)
        tie, _, _ = O.make_inputs(cfg, 50, 1, seed=1, dtype=torch.bfloat16)   # two query blocks, two key tiles
        lg, hid = p.engine.prefill(tie.cuda(), n_left_pad=0)
        torch.cuda.synchronize()
        print("OK prefill P=50", tuple(lg.shape), bool(torch.isfinite(lg.float()).all()), flush=True)
    else:
        print("prefill weights not set for this geometry", flush=True)
