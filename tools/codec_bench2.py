import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "faster-qwen3-tts_b200")]
import torch
from faster_qwen3_tts.codec import build_codec
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
st = build_codec(dtype=torch.bfloat16, device="cuda", seed=1, backend="engine")
st_t = build_codec(dtype=torch.bfloat16, device="cuda", seed=1, backend="torch")
def ev(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
for T in (33, 182):
    codes = torch.randint(0, 2048, (1, T, 16), device="cuda")
    a = st.decode({"audio_codes": codes})[0][0]; b = st_t.decode({"audio_codes": codes})[0][0]
    ms_e = ev(lambda: st.decode({"audio_codes": codes}))
    ms_t = ev(lambda: st_t.decode({"audio_codes": codes}))
    # stack only: front output -> kernels
    with torch.inference_mode():
        x = st._front_graphed(codes.transpose(1, 2).contiguous())[0].to(torch.bfloat16).contiguous().clone()
    import ctypes as C
    pcm = torch.empty(T * 1920, dtype=torch.float32, device="cuda")
    def stack():
        st._lib.fq3_codec_decode(st._h, C.c_void_p(x.data_ptr()), x.shape[1], C.c_void_p(pcm.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    ms_s = ev(stack)
    fl = st.flops(T)
    print(json.dumps({"T": T, "engine_ms": ms_e, "torch_ms": ms_t, "stack_ms": ms_s, "stack_gflop": fl / 1e9,
                      "stack_tflops": fl / ms_s / 1e9, "frac_of_bf16_peak": fl / ms_s / 1e9 / peaks.get("bf16_tflops", 1680.0),
                      "max_abs_diff_vs_torch": (a - b).abs().max().item()}))
