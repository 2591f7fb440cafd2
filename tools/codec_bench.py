import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "faster-qwen3-tts_b200")]
import torch
from faster_qwen3_tts.codec import build_codec
st = build_codec(dtype=torch.bfloat16, device="cuda", seed=1)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
for T in (33, 182):
    codes = torch.randint(0, 2048, (1, T, 16), device="cuda")
    ms = t(lambda: st.decode({"audio_codes": codes}))
    # graph
    g = torch.cuda.CUDAGraph()
    with torch.inference_mode():
        x = codes.transpose(1, 2).contiguous()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2): st.decoder(x)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            y = st.decoder(x)
    gms = t(lambda: g.replay())
    print(json.dumps({"T": T, "eager_ms": ms, "graph_ms": gms}))
# per-part timing at T=33 (eager, synchronised): pre-transformer+upsample vs decoder blocks
dec = st.decoder
with torch.inference_mode():
    T=33; codes = torch.randint(0, 2048, (1, 16, T), device="cuda")
    import torch.nn.functional as F
    def front():
        c = dec.config
        off = (torch.arange(16, device="cuda") * c.codebook_size).view(1, 16, 1)
        x = dec.code_embedding(codes + off).mean(1)
        hd = c.hidden_size // c.num_attention_heads
        inv = 1.0 / (c.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32, device="cuda") / hd))
        fr = torch.arange(T, dtype=torch.float32, device="cuda")[:, None] * inv[None]
        emb = torch.cat((fr, fr), dim=-1)
        cos, sin = emb.cos().to(x.dtype)[None, None], emb.sin().to(x.dtype)[None, None]
        i = torch.arange(T, device="cuda")
        allowed = (i[None, :] <= i[:, None]) & (i[None, :] > i[:, None] - c.sliding_window)
        for l in dec.layers: x = l(x, cos, sin, allowed)
        x = dec.norm(x).transpose(1, 2)
        for up, nx in dec.upsample: x = nx(up(x))
        return dec.conv_in(x)
    x0 = front()
    print(json.dumps({"front_ms": t(front)}))
    cur = x0
    for bi, b in enumerate(dec.blocks):
        inp = cur
        print(json.dumps({"block": bi, "in_shape": list(inp.shape), "ms": t(lambda: b(inp))}))
        cur = b(inp)
