#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native Qwen3-TTS decode engine.

Metric (BASELINE.json): xRealTime (RTF = audio seconds / wall seconds) and p50 TTFA, Qwen3-TTS-12Hz-1.7B streaming,
chunk_size=8, on 1/2/4/8 B200 (independent replicas, no collective on this path).

A "step" is one streaming request of SURVEY.md section 8(d) config 3: prompt P=232 (ICL-shaped), 128 frames (10.24 s of
audio) generated in chunks of 8, every chunk decoded to 24 kHz PCM with the reference's two-phase window policy.
min_new_tokens = max_new_tokens pins the work per step (EOS cannot cut a run short).  Weights are random-init at the
real 1.7B geometry, inputs synthetic (no checkpoint / tokenizer exists offline).

  value  : RTF with prompt embeddings already resident in HBM, CUDA-event timed, max over ranks
  e2e    : RTF through the public API (FasterQwen3TTS.generate_voice_clone_streaming) with prompt embeddings in
           pinned HOST memory (H2D inside the timed region) and every PCM chunk read back to the host (D2H)
  roofline: persistent decode kernel, algorithmic bytes per launch (SURVEY.md 8(d) B_alg) / CUDA-event launch time
            against the MEASURED HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU oracle (torch eager, dynamic KV; the reference refuses non-CUDA devices,
            model.py:181-182, and its arithmetic lives in absent third-party packages, so the oracle port IS its
            CPU path) on a bounded sample of the same workload, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "faster-qwen3-tts_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line (NCCL is only used for the timing barrier)
import torch  # noqa: E402

FRAME_S = 0.08  # 1920 samples @ 24 kHz (ggml_backend.py:22)
METRIC = "xRealTime (RTF) Qwen3-TTS-1.7B streaming chunk_size=8 (p50 TTFA in config)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", default="1.7B", choices=["1.7B", "0.6B"])
    ap.add_argument("--prompt", type=int, default=232)
    ap.add_argument("--trailing", type=int, default=25)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--chunk", type=int, default=8)
    ap.add_argument("--ref-frames", type=int, default=174, help="ICL reference codes prepended in codec Phase 1")
    ap.add_argument("--no-codec", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--num-ctas", type=int, default=0)
    ap.add_argument("--cpu-frames", type=int, default=64)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle on host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_oracle_run(args, frames: int):
    """prefill(P) + `frames` decode frames on the CPU oracle (fp32 eager -- torch CPU bf16 GEMV is ~50x slower than
    fp32 on this host, measured 479 s for the same sample -- dynamic KV, all host threads).
    Returns (rtf, seconds, threads, description)."""
    from oracle import qwen3_tts_oracle as O
    # torch-eager GEMV chains stop scaling (and collapse under OpenMP oversubscription: 479 s for this sample with 128
    # threads on the GPU box vs 3 s with 8 threads) -- use at most 16 threads and report the number used
    nthreads = min(os.cpu_count() or 1, 16)
    torch.set_num_threads(nthreads)
    cfg = O.cfg_1p7b() if args.size == "1.7B" else O.cfg_0p6b()
    # cheap deterministic weights (values do not matter for timing; shapes/dtype do)
    t0 = time.time()
    W = {}
    g = torch.Generator().manual_seed(0)

    def fill(shape, std):
        return (torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g))

    def stack(prefix, c):
        qd, kd = c.num_attention_heads * 128, c.num_key_value_heads * 128
        for li in range(c.num_hidden_layers):
            p = f"{prefix}.layers.{li}."
            for n, sh in (("self_attn.q_proj", (qd, c.hidden_size)), ("self_attn.k_proj", (kd, c.hidden_size)),
                          ("self_attn.v_proj", (kd, c.hidden_size)), ("self_attn.o_proj", (c.hidden_size, qd)),
                          ("mlp.gate_proj", (c.intermediate_size, c.hidden_size)),
                          ("mlp.up_proj", (c.intermediate_size, c.hidden_size)),
                          ("mlp.down_proj", (c.hidden_size, c.intermediate_size))):
                W[p + n + ".weight"] = fill(sh, 0.02)
            for n, k in (("input_layernorm", c.hidden_size), ("post_attention_layernorm", c.hidden_size),
                         ("self_attn.q_norm", 128), ("self_attn.k_norm", 128)):
                W[p + n + ".weight"] = torch.ones(k, dtype=torch.float32)
        W[prefix + ".norm.weight"] = torch.ones(c.hidden_size, dtype=torch.float32)

    Ht, Hp = cfg.talker.hidden_size, cfg.predictor.hidden_size
    stack("talker.model", cfg.talker)
    W["talker.model.codec_embedding.weight"] = fill((cfg.talker.vocab_size, Ht), 1.0)
    W["talker.codec_head.weight"] = fill((cfg.talker.vocab_size, Ht), 0.08)
    stack("talker.code_predictor.model", cfg.predictor)
    for i in range(15):
        W[f"talker.code_predictor.model.codec_embedding.{i}.weight"] = fill((cfg.predictor.vocab_size, Ht), 1.0)
        W[f"talker.code_predictor.lm_head.{i}.weight"] = fill((cfg.predictor.vocab_size, Hp), 0.08)
    if cfg.has_mtp_projection:
        W["talker.code_predictor.small_to_mtp_projection.weight"] = fill((Hp, Ht), 0.02)
        W["talker.code_predictor.small_to_mtp_projection.bias"] = fill((Hp,), 0.02)
    om = O.OracleModel(cfg, W, max_pos=args.prompt + frames + 8)
    tie, tth, tpe = O.make_inputs(cfg, args.prompt, args.trailing, seed=0, dtype=torch.float32)
    import numpy as np
    u = np.random.default_rng(0).random((frames + 1, 16), dtype=np.float32)
    t_build = time.time() - t0
    with torch.inference_mode():
        t1 = time.time()
        codes = O.generate(om, tie, tth, tpe, max_new_tokens=frames, min_new_tokens=frames, uniforms=u,
                           max_seq_len=2048)
        dt = time.time() - t1
    n = int(codes.shape[0])
    desc = (f"CPU oracle (torch eager fp32, dynamic KV): prefill P={args.prompt} + {n} frames of the {args.size} "
            f"workload, {nthreads} threads, {dt:.1f}s (weights built in {t_build:.0f}s, untimed); no codec decode")
    return n * FRAME_S / dt, dt, nthreads, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        rtf, dt, nth, desc = cpu_oracle_run(args, args.cpu_frames)
        if i >= args.warmup:
            vals.append((rtf, dt))
        if i == 0 and args.warmup > 0 and dt > 60:  # keep the whole run within minutes
            args.warmup = 0
            vals.append((rtf, dt))
            break
    v = statistics.mean(x[0] for x in vals)
    ms = statistics.mean(x[1] for x in vals) * 1000
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "x realtime", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, sample=f"prefill + {args.cpu_frames} frames per step"),
        "cpu_baseline": {"value": v, "unit": "x realtime", "cores": nth, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, **extra):
    c = {"workload": f"Qwen3-TTS-12Hz-{args.size} streaming voice-clone (SURVEY 8(d) config 3): prompt P={args.prompt}, "
                     f"trailing text {args.trailing}, {args.frames} frames, chunk_size={args.chunk}, "
                     f"T=0.9 top_k=50 top_p=1.0 penalty=1.05, min_new_tokens=max_new_tokens (fixed work)",
         "batch_per_gpu": 1, "parallelism": f"replicas x{args.gpus} (no collective)",
         "l2_policy": "per-step weight stream (3.2 GB tape) exceeds the 126 MB L2; no explicit flush needed"}
    c.update(extra)
    return c


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from faster_qwen3_tts import synthetic
    from faster_qwen3_tts.model import FasterQwen3TTS
    from faster_qwen3_tts.streaming import fast_generate_streaming

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.make_config(args.size)
    model = FasterQwen3TTS.from_synthetic(args.size, device=f"cuda:{local}", dtype=torch.bfloat16, max_seq_len=2048,
                                          seed=1234 + rank, num_ctas=args.num_ctas, with_codec=not args.no_codec)
    eng = model.engine
    eng.time_kernels = True
    talker = model.model.model.talker
    tcfg = cfg.talker_config
    tie, tam, tth, tpe = synthetic.make_prompt(cfg, args.prompt, args.trailing, seed=rank, dtype=torch.bfloat16, device=dev)
    host_prompt = [t.cpu().pin_memory() for t in (tie, tth, tpe)]
    ref_codes = None
    if args.ref_frames > 0:
        ref_codes = torch.randint(0, 2048, (args.ref_frames, 16), device=dev)
    kw = dict(max_new_tokens=args.frames, min_new_tokens=args.frames, chunk_size=args.chunk)
    launches0 = [0]
    chunk_ms, ttfa_ms = [], []

    def step_resident(timed: bool):
        """prompt resident in HBM; codes -> PCM per chunk on device.  Returns frames."""
        torch.manual_seed(rank * 1000 + len(ttfa_ms))
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        first = None
        n = 0
        gen = model.stream_from_embeds(tie, tam, tth, tpe, ref_codes=ref_codes, to_host=False, **kw)
        for pcm, sr, t in gen:
            if first is None:
                first = torch.cuda.Event(enable_timing=True)
                first.record()
            n += t["chunk_steps"]
            if timed and "kernel_ms" in t:
                chunk_ms.append(t["kernel_ms"])
        if timed and first is not None:
            first.synchronize()
            ttfa_ms.append(e0.elapsed_time(first))
        return n

    def step_e2e():
        """public API with HOST buffers: H2D of the prompt and D2H of every PCM chunk inside the timed region."""
        h2d = d2h = 0
        d = [t.to(dev, non_blocking=True) for t in host_prompt]
        h2d += sum(t.numel() * t.element_size() for t in host_prompt)
        n = 0
        t_first = None
        t0 = time.perf_counter()
        for pcm, sr, t in model.stream_from_embeds(d[0], tam, d[1], d[2], ref_codes=ref_codes, to_host=True, **kw):
            if t_first is None:
                t_first = time.perf_counter() - t0
            d2h += pcm.nbytes
            n += t["chunk_steps"]
        return n, h2d, d2h, t_first

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident(False)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    barrier()
    l0 = eng.launch_count + model.codec_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    frames = 0
    for _ in range(args.steps):
        frames += step_resident(True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count + model.codec_launches() - l0
    clk = clocks.stop() if rank == 0 else None
    # e2e leg
    barrier()
    t0 = time.perf_counter()
    e_frames = h2d = d2h = 0
    e_ttfa = []
    for _ in range(args.steps):
        n, a, b, tf = step_e2e()
        e_frames += n
        h2d, d2h = a, b
        e_ttfa.append(tf * 1000)
    torch.cuda.synchronize()
    e_s = time.perf_counter() - t0
    from faster_qwen3_tts.replicas import aggregate
    (ms, e_ms), (frames, e_frames) = aggregate([ms, e_s * 1000], [frames, e_frames], device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = frames * FRAME_S / (ms / 1000)
    e2e = e_frames * FRAME_S / (e_ms / 1000)
    # ---- roofline of the persistent decode kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    t_bytes, p_bytes = eng.tape_bytes()
    esz = 2
    Lt, nKV = tcfg.num_hidden_layers, tcfg.num_key_value_heads
    pbar = args.prompt + (args.frames - 1) / 2.0
    kv_bytes = Lt * 2 * nKV * 128 * esz * (pbar + 1)
    # B_alg: every distinct weight byte once per frame; B_stream: bytes the tape actually streams
    from faster_qwen3_tts.weights import stack_config
    pc = stack_config(cfg.code_predictor_config)
    pred_layers = pc["num_hidden_layers"] * (2 * pc["hidden_size"] * (pc["num_attention_heads"] + pc["num_key_value_heads"]) * 128
                                             + 3 * pc["hidden_size"] * pc["intermediate_size"]) * esz
    pred_heads = 15 * pc["vocab_size"] * pc["hidden_size"] * esz
    mtp = (tcfg.hidden_size * pc["hidden_size"] * esz) if cfg.has_mtp else 0
    b_alg = t_bytes + kv_bytes + pred_layers + pred_heads + mtp
    b_stream = t_bytes + kv_bytes + p_bytes
    k_ms = statistics.mean(chunk_ms) if chunk_ms else None
    traffic = None
    traffic_file = next((f for f in ("r1c_decode_kernel_ncu.csv", "r1b_decode_kernel_ncu.csv")
                         if os.path.exists(os.path.join(ROOT, "profiles", f))), "r1b_decode_kernel_ncu.csv")
    try:  # DRAM bytes of one launch from the committed ncu capture of this kernel (profiles/)
        for line in open(os.path.join(ROOT, "profiles", traffic_file)):
            f = line.strip().split(",")
            if len(f) == 4 and f[0] == "0" and f[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                traffic = (traffic or 0.0) + float(f[3]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[2]]
    except Exception:
        traffic = None
    # the talker step alone (the kernel the north_star's HBM target is stated for): MODE_TALKER_STEP launches at the
    # run's mean position, bytes = talker layer weights + KV(p) (no head), timed live with CUDA events
    talker = None
    try:
        ppos = int(pbar)
        xh = torch.randn(tcfg.hidden_size, device=dev).to(torch.bfloat16)
        for _ in range(3):
            eng.talker_step(xh, ppos)
        torch.cuda.synchronize()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(20):
            eng.talker_step(xh, ppos)
        t1e.record(); t1e.synchronize()
        ts_ms = t0e.elapsed_time(t1e) / 20
        ts_bytes = t_bytes - tcfg.vocab_size * tcfg.hidden_size * esz + Lt * 2 * nKV * 128 * esz * (ppos + 1)
        talker = {"position": ppos, "ms": ts_ms, "bytes": ts_bytes, "achieved": ts_bytes / (ts_ms / 1000) / 1e9,
                  "frac": ts_bytes / (ts_ms / 1000) / 1e9 / peak,
                  "note": "one launch per step here (launch + pipeline fill included); inside the fused loop the step is shorter"}
    except Exception as ex:  # never let the extra measurement break the bench line
        talker = {"error": str(ex)[:120]}
    roof = None
    if k_ms:
        ach = b_alg * args.chunk / (k_ms / 1000) / 1e9
        roof = {"bound": "hbm", "kernel": "fq3_decode_kernel<bf16> (one launch = one 8-frame chunk)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": "ncu --set full capture of one 8-frame launch, profiles/" + traffic_file,
                "peak_source": peak_kind, "alg_bytes_per_frame": b_alg, "launch_ms": k_ms,
                "streamed_bytes_per_frame": b_stream, "streamed_frac": b_stream * args.chunk / (k_ms / 1000) / 1e9 / peak,
                "ms_per_frame": k_ms / args.chunk, "talker_step": talker}
    out = {
        "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, ttfa_ms_p50=statistics.median(ttfa_ms) if ttfa_ms else None,
                                  ttfa_ms_e2e_p50=statistics.median(e_ttfa) if e_ttfa else None,
                                  codec=not args.no_codec, ctas=eng.num_ctas, ref_frames=args.ref_frames),
        "e2e": {"value": e2e, "unit": "x realtime", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            rtf, dt, nth, desc = cpu_oracle_run(args, args.cpu_frames)
            out["cpu_baseline"] = {"value": rtf, "unit": "x realtime", "cores": nth, "kind": "port", "sample": desc}
        except Exception as ex:  # the bench line must still print
            out["cpu_baseline"] = {"value": None, "unit": "x realtime", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"failed: {ex!r}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
