"""ctypes binding of the B200-native decode engine (include/fq3_engine.h).

Python stays a thin host: every tensor crossing this boundary is a torch CUDA tensor whose ``data_ptr()`` is
handed to the C ABI; all arithmetic of the decode loop happens inside ``libfq3_engine.so`` (hand-written
sm_100a CUDA, csrc/).  There is NO fallback: if the shared library is missing or no CUDA device is present the
constructors raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
LIB_PATH = os.path.join(CSRC, "libfq3_engine.so")
INCLUDE = os.path.normpath(os.path.join(_HERE, "..", "..", "include"))

FQ3_F32, FQ3_BF16 = 0, 1
FINISH_NAMES = {0: "running", 1: "max_new_tokens", 2: "eos", 3: "max_seq_len"}


class StackConfig(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("intermediate_size", C.c_int32), ("num_hidden_layers", C.c_int32),
                ("num_attention_heads", C.c_int32), ("num_key_value_heads", C.c_int32), ("vocab_size", C.c_int32),
                ("rms_norm_eps", C.c_float)]


class Config(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("device", C.c_int32), ("max_seq_len", C.c_int32),
                ("num_code_groups", C.c_int32), ("codec_eos_token_id", C.c_int32),
                ("has_mtp_projection", C.c_int32), ("num_ctas", C.c_int32), ("rope_positions", C.c_int32),
                ("talker", StackConfig), ("predictor", StackConfig), ("max_batch", C.c_int32)]


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dev_ptr", C.c_void_p), ("numel", C.c_int64)]


class Sampling(C.Structure):
    _fields_ = [("do_sample", C.c_int32), ("top_k", C.c_int32), ("temperature", C.c_float), ("top_p", C.c_float),
                ("repetition_penalty", C.c_float)]


class Request(C.Structure):
    _fields_ = [("first_token", C.c_int32), ("prefill_len", C.c_int32), ("gen_step", C.c_int32),
                ("rope_delta", C.c_int32), ("n_left_pad", C.c_int32), ("max_new_tokens", C.c_int32),
                ("min_new_tokens", C.c_int32), ("trailing_len", C.c_int32)]


class ChunkResult(C.Structure):
    _fields_ = [("frames_emitted", C.c_int32), ("finished", C.c_int32), ("total_frames", C.c_int32),
                ("next_token", C.c_int32)]


EXPORTS = [
    "fq3_engine_create", "fq3_engine_load_weights", "fq3_engine_destroy", "fq3_import_kv", "fq3_export_kv",
    "fq3_set_generation_state", "fq3_talker_step", "fq3_predictor_run", "fq3_sample_logits", "fq3_begin_request",
    "fq3_decode_chunk", "fq3_get_past_hidden", "fq3_debug_enable", "fq3_debug_read", "fq3_tape_bytes",
    "fq3_num_ctas", "fq3_launch_count", "fq3_last_error", "fq3_version", "fq3_barrier_test",
    "fq3_engine_set_prefill_weights", "fq3_prefill", "fq3_set_gemm_backend", "fq3_max_batch", "fq3_debug_gemv",
    "fq3_codec_create", "fq3_codec_load_weights", "fq3_codec_decode", "fq3_codec_decode_batch", "fq3_codec_flops",
    "fq3_codec_load_frontend", "fq3_codec_decode_codes", "fq3_codec_frontend_flops",
    "fq3_codec_stream_create", "fq3_codec_stream_reset", "fq3_codec_stream_destroy", "fq3_codec_stream_frames",
    "fq3_codec_stream_decode", "fq3_codec_stream_copy", "fq3_codec_launch_count",
    "fq3_codec_destroy", "fq3_codec_last_error",
]


def build_extension(verbose: bool = False) -> str:
    """Compile csrc/*.cu into libfq3_engine.so for sm_100a (cross-compiles without a GPU)."""
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-I", INCLUDE, "-o", LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU / PyTorch fallback for the decode path.")
    lib = C.CDLL(LIB_PATH)
    lib.fq3_last_error.restype = C.c_char_p
    lib.fq3_version.restype = C.c_char_p
    lib.fq3_launch_count.restype = C.c_int64
    lib.fq3_launch_count.argtypes = [C.c_void_p]
    lib.fq3_engine_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.fq3_engine_load_weights.argtypes = [C.c_void_p, C.POINTER(Tensor), C.c_int32, C.c_void_p]
    lib.fq3_engine_destroy.argtypes = [C.c_void_p]
    lib.fq3_engine_destroy.restype = None
    lib.fq3_import_kv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.fq3_export_kv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.fq3_set_generation_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.fq3_talker_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.fq3_predictor_run.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(Sampling), C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    lib.fq3_sample_logits.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Sampling), C.c_float, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.fq3_begin_request.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Request), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(Sampling), C.POINTER(Sampling), C.c_void_p]
    lib.fq3_decode_chunk.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p,
                                     C.POINTER(ChunkResult), C.c_void_p]
    lib.fq3_get_past_hidden.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.fq3_max_batch.argtypes = [C.c_void_p]
    lib.fq3_debug_gemv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fq3_debug_enable.argtypes = [C.c_void_p, C.c_int32]
    lib.fq3_debug_read.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.fq3_tape_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.fq3_num_ctas.argtypes = [C.c_void_p]
    lib.fq3_barrier_test.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.fq3_set_gemm_backend.argtypes = [C.c_int32]
    lib.fq3_engine_set_prefill_weights.argtypes = [C.c_void_p, C.POINTER(Tensor), C.c_int32]
    lib.fq3_prefill.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p]
    lib.fq3_codec_create.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]
    lib.fq3_codec_destroy.argtypes = [C.c_void_p]
    lib.fq3_codec_destroy.restype = None
    lib.fq3_codec_load_weights.argtypes = [C.c_void_p, C.POINTER(Tensor), C.c_int32, C.c_void_p]
    lib.fq3_codec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.fq3_codec_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.fq3_codec_load_frontend.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_float), C.c_int32,
                                            C.POINTER(Tensor), C.c_int32, C.c_void_p]
    lib.fq3_codec_decode_codes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.fq3_codec_stream_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.fq3_codec_stream_reset.argtypes = [C.c_void_p, C.c_void_p]
    lib.fq3_codec_stream_destroy.argtypes = [C.c_void_p]
    lib.fq3_codec_stream_destroy.restype = None
    lib.fq3_codec_stream_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fq3_codec_stream_frames.argtypes = [C.c_void_p]
    lib.fq3_codec_stream_frames.restype = C.c_int64
    lib.fq3_codec_stream_decode.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                            C.c_void_p]
    lib.fq3_codec_frontend_flops.argtypes = [C.c_void_p, C.c_int32]
    lib.fq3_codec_frontend_flops.restype = C.c_double
    lib.fq3_codec_flops.argtypes = [C.c_void_p, C.c_int32]
    lib.fq3_codec_flops.restype = C.c_double
    lib.fq3_codec_launch_count.argtypes = [C.c_void_p]
    lib.fq3_codec_launch_count.restype = C.c_int64
    lib.fq3_codec_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def set_gemm_backend(name: str):
    """Dense layers of K3 (prefill) and K4 (codec): 'tcgen05' (default: one 128x96 tile per CTA, 2-4 CTAs per SM),
    'tcgen05_persistent' (tile loop with a double-buffered TMEM accumulator; measured slower on these shapes) or 'mma'
    (mma.sync kernel) -- the last two for A/B runs."""
    load_library().fq3_set_gemm_backend({"tcgen05": 0, "mma": 1, "tcgen05_persistent": 2}[name])


class EngineError(RuntimeError):
    pass


def _check(lib, rc: int):
    if rc != 0:
        msg = lib.fq3_last_error().decode()
        if rc == -4:
            raise RuntimeError(msg)  # same type/message as talker_graph.py:163-167
        raise EngineError(f"fq3 error {rc}: {msg}")


@dataclass
class SamplingParams:
    do_sample: bool = True
    top_k: int = 50
    temperature: float = 0.9
    top_p: float = 1.0
    repetition_penalty: float = 1.0

    def c(self) -> Sampling:
        return Sampling(int(bool(self.do_sample)), int(self.top_k), float(self.temperature), float(self.top_p),
                        float(self.repetition_penalty))


class Engine:
    """One engine per device: packed weights, KV caches and the persistent decode kernel."""

    def __init__(self, *, talker: dict, predictor: dict, dtype: torch.dtype, device="cuda", max_seq_len: int = 2048,
                 num_code_groups: int = 16, codec_eos_token_id: int = 2150, has_mtp_projection: bool = True,
                 num_ctas: int = 0, rope_positions: Optional[int] = None, max_batch: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("fq3 engine needs a CUDA device (sm_100a); no CPU fallback exists")
        self.lib = load_library()
        dev = torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("engine dtype must be torch.float32 or torch.bfloat16")
        self.dtype = dtype
        self.max_seq_len = int(max_seq_len)
        self.talker_cfg, self.pred_cfg = dict(talker), dict(predictor)
        self.num_code_groups = num_code_groups
        self.eos = codec_eos_token_id
        self.rope_positions = int(rope_positions or (self.max_seq_len + 64))

        def sc(d):
            return StackConfig(d["hidden_size"], d["intermediate_size"], d["num_hidden_layers"],
                               d["num_attention_heads"], d["num_key_value_heads"], d["vocab_size"],
                               float(d.get("rms_norm_eps", 1e-6)))

        cfg = Config(FQ3_BF16 if dtype == torch.bfloat16 else FQ3_F32, self.device.index, self.max_seq_len,
                     num_code_groups, codec_eos_token_id, int(bool(has_mtp_projection)), int(num_ctas),
                     self.rope_positions, sc(talker), sc(predictor), int(max_batch))
        self.max_batch = int(max_batch)
        h = C.c_void_p()
        _check(self.lib, self.lib.fq3_engine_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.H = talker["hidden_size"]
        self._keep = {}  # slot -> tensors borrowed by the engine for the duration of that slot's request
        self.loaded = False
        self.has_prefill = False
        self._prefill_keep = None
        self.time_kernels = False   # bench: CUDA-event time of every decode_chunk launch
        self.last_kernel_ms = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.fq3_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _t(self, x: torch.Tensor, dtype=None) -> torch.Tensor:
        x = x.to(device=self.device, dtype=dtype or self.dtype)
        return x if x.is_contiguous() else x.contiguous()

    # -- weights ----------------------------------------------------------------------------------------
    def load_weights(self, tensors: Dict[str, torch.Tensor]):
        keep, arr = [], (Tensor * len(tensors))()
        for i, (name, t) in enumerate(tensors.items()):
            want = torch.float32 if name.split(".")[-1] in ("cos", "sin") else self.dtype
            t = self._t(t, want)
            keep.append(t)
            arr[i] = Tensor(name.encode(), t.data_ptr(), t.numel())
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.fq3_engine_load_weights(self.h, arr, len(tensors), self._stream()))
        self.loaded = True

    # -- K3 prefill -------------------------------------------------------------------------------------
    def set_prefill_weights(self, tensors: Dict[str, torch.Tensor]):
        """Row-major bf16 weights borrowed by the hand-written prefill (kept alive here)."""
        keep = {k: self._t(v) for k, v in tensors.items()}
        arr = (Tensor * len(keep))()
        for i, (k, v) in enumerate(keep.items()):
            arr[i] = Tensor(k.encode(), v.data_ptr(), v.numel())
        _check(self.lib, self.lib.fq3_engine_set_prefill_weights(self.h, arr, len(keep)))
        self._prefill_keep = keep
        self.has_prefill = True

    def prefill(self, embeds: torch.Tensor, n_left_pad: int = 0, slot: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """embeds [P,H] -> (logits [V], past_hidden [H]); cache rows [0,P) of request slot `slot` are written."""
        x = self._t(embeds.reshape(-1, self.H))
        logits = torch.empty(self.talker_cfg["vocab_size"], dtype=self.dtype, device=self.device)
        hidden = torch.empty(self.H, dtype=self.dtype, device=self.device)
        _check(self.lib, self.lib.fq3_prefill(self.h, int(slot), x.data_ptr(), x.shape[0], int(n_left_pad), logits.data_ptr(),
                                              hidden.data_ptr(), self._stream()))
        return logits, hidden

    # -- duck-type path ----------------------------------------------------------------------------------
    def import_kv(self, layer: int, k: torch.Tensor, v: torch.Tensor, slot: int = 0):
        """k, v: [1, n_kv, P, 128] (HF cache layout) or [n_kv, P, 128]."""
        k, v = self._t(k.reshape(-1, k.shape[-2], k.shape[-1])), self._t(v.reshape(-1, v.shape[-2], v.shape[-1]))
        _check(self.lib, self.lib.fq3_import_kv(self.h, int(slot), layer, k.data_ptr(), v.data_ptr(), k.shape[1],
                                                self._stream()))
        return k.shape[1]

    def export_kv(self, layer: int, P: int, slot: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """cache rows [0,P) of one layer as (k, v) [n_kv, P, 128]"""
        nkv = self.talker_cfg["num_key_value_heads"]
        k = torch.empty(nkv, P, 128, dtype=self.dtype, device=self.device)
        v = torch.empty_like(k)
        _check(self.lib, self.lib.fq3_export_kv(self.h, int(slot), int(layer), k.data_ptr(), v.data_ptr(), int(P),
                                                self._stream()))
        return k, v

    def set_generation_state(self, n_left_pad: int, rope_delta: int, slot: int = 0):
        _check(self.lib, self.lib.fq3_set_generation_state(self.h, int(slot), int(n_left_pad), int(rope_delta)))

    def talker_step(self, embeds: torch.Tensor, position: int, out: Optional[torch.Tensor] = None,
                    slot: int = 0) -> torch.Tensor:
        x = self._t(embeds.reshape(-1))
        if out is None:
            out = torch.empty(self.H, dtype=self.dtype, device=self.device)
        _check(self.lib, self.lib.fq3_talker_step(self.h, int(slot), x.data_ptr(), int(position), out.data_ptr(),
                                                  self._stream()))
        return out

    def predictor_run(self, pred_input: torch.Tensor, sp: SamplingParams,
                      uniforms: Optional[torch.Tensor] = None, slot: int = 0) -> torch.Tensor:
        x = self._t(pred_input.reshape(2, -1))
        out = torch.empty(self.num_code_groups - 1, dtype=torch.long, device=self.device)
        u = None
        if sp.do_sample:
            if uniforms is None:
                uniforms = torch.rand(self.num_code_groups - 1, device=self.device)
            u = self._t(uniforms, torch.float32)
        s = sp.c()
        _check(self.lib, self.lib.fq3_predictor_run(self.h, int(slot), x.data_ptr(), C.byref(s), u.data_ptr() if u is not None else None,
                                                     out.data_ptr(), self._stream()))
        return out

    def sample_logits(self, logits: torch.Tensor, sp: SamplingParams, u: float = 0.0,
                      history: Optional[torch.Tensor] = None, suppress_special: bool = False, eos_id: int = -1,
                      suppress_eos: bool = False) -> torch.Tensor:
        lg = self._t(logits.reshape(-1))
        out = torch.empty(1, dtype=torch.long, device=self.device)
        h = self._t(history.reshape(-1), torch.long) if history is not None and history.numel() else None
        s = sp.c()
        _check(self.lib, self.lib.fq3_sample_logits(self.h, lg.data_ptr(), lg.numel(), C.byref(s), float(u),
                                                     h.data_ptr() if h is not None else None,
                                                     h.numel() if h is not None else 0, int(suppress_special), int(eos_id),
                                                     int(suppress_eos), out.data_ptr(), self._stream()))
        return out

    # -- fused path ----------------------------------------------------------------------------------------
    def begin_request(self, *, first_token: int, prefill_len: int, gen_step: int, past_hidden: torch.Tensor,
                      trailing_text: torch.Tensor, tts_pad: torch.Tensor, max_new_tokens: int, min_new_tokens: int,
                      sp_talker: SamplingParams, sp_predictor: SamplingParams, uniforms: Optional[torch.Tensor],
                      rope_delta: int = 0, n_left_pad: int = 0, slot: int = 0):
        ph = self._t(past_hidden.reshape(-1))
        tt = self._t(trailing_text.reshape(-1, self.H)) if trailing_text is not None and trailing_text.numel() else None
        tp = self._t(tts_pad.reshape(-1))
        need_u = sp_talker.do_sample or sp_predictor.do_sample
        if need_u and uniforms is None:
            uniforms = torch.rand(max_new_tokens + 1, 16, device=self.device)
        u = self._t(uniforms, torch.float32) if uniforms is not None else None
        if u is not None and u.numel() < (max_new_tokens + 1) * 16:
            raise ValueError("uniforms must have (max_new_tokens + 1) * 16 elements")
        self._keep[int(slot)] = dict(ph=ph, tt=tt, tp=tp, u=u)
        rq = Request(int(first_token), int(prefill_len), int(gen_step), int(rope_delta), int(n_left_pad),
                     int(max_new_tokens), int(min_new_tokens), 0 if tt is None else tt.shape[0])
        st, spd = sp_talker.c(), sp_predictor.c()
        _check(self.lib, self.lib.fq3_begin_request(
            self.h, int(slot), C.byref(rq), ph.data_ptr(), tt.data_ptr() if tt is not None else None, tp.data_ptr(),
            u.data_ptr() if u is not None else None, C.byref(st), C.byref(spd), self._stream()))

    def decode_chunk(self, n_frames: int, out: Optional[torch.Tensor] = None, slot: int = 0) -> Tuple[torch.Tensor, ChunkResult]:
        """Single-sequence launch on one slot: (codes [frames_emitted,16], result)."""
        if out is None:
            out = torch.empty(n_frames, 16, dtype=torch.long, device=self.device)
        res = ChunkResult()
        sl = (C.c_int32 * 1)(int(slot))
        if self.time_kernels:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _check(self.lib, self.lib.fq3_decode_chunk(self.h, sl, 1, int(n_frames), out.data_ptr(), C.byref(res), self._stream()))
        if self.time_kernels:
            e1.record()
            e1.synchronize()
            self.last_kernel_ms = e0.elapsed_time(e1)
        return out[: res.frames_emitted], res

    def decode_chunk_batch(self, slots, n_frames: int, out: Optional[torch.Tensor] = None):
        """All listed slots advance up to n_frames frames in ONE launch sharing every pass over the weights.
        Returns (codes [n_slots, n_frames, 16] -- row j valid up to results[j].frames_emitted --, [ChunkResult])."""
        slots = [int(x) for x in slots]
        n = len(slots)
        if out is None:
            out = torch.empty(n, n_frames, 16, dtype=torch.long, device=self.device)
        res = (ChunkResult * n)()
        sl = (C.c_int32 * n)(*slots)
        if self.time_kernels:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _check(self.lib, self.lib.fq3_decode_chunk(self.h, sl, n, int(n_frames), out.data_ptr(), res, self._stream()))
        if self.time_kernels:
            e1.record()
            e1.synchronize()
            self.last_kernel_ms = e0.elapsed_time(e1)
        return out, list(res)

    def past_hidden(self, slot: int = 0) -> torch.Tensor:
        out = torch.empty(self.H, dtype=self.dtype, device=self.device)
        _check(self.lib, self.lib.fq3_get_past_hidden(self.h, int(slot), out.data_ptr(), self._stream()))
        return out

    def debug_gemv(self, stack: int, layer: int, which: int, x: torch.Tensor) -> torch.Tensor:
        """One batched GEMV over a weight segment (numerics probe): x [ncols,K] -> [ncols, rows] fp32
        (which == 2: model dtype [ncols, I] = silu(gate) * up)."""
        x = self._t(x)
        d = self.talker_cfg if stack == 0 else self.pred_cfg
        H, I = d["hidden_size"], d["intermediate_size"]
        rows = {0: (d["num_attention_heads"] + 2 * d["num_key_value_heads"]) * 128, 1: H, 2: I, 3: H, 4: d["vocab_size"]}[which]
        out = torch.empty(x.shape[0], rows, dtype=self.dtype if which == 2 else torch.float32, device=self.device)
        _check(self.lib, self.lib.fq3_debug_gemv(self.h, int(stack), int(layer), int(which), x.shape[0], x.data_ptr(),
                                                 out.data_ptr(), self._stream()))
        return out

    # -- introspection ----------------------------------------------------------------------------------------
    def debug_enable(self, on):
        """bit 0: dump per-layer intermediates; bit 1: clock64 probes of CTA 0 (tools/microbench.py)."""
        _check(self.lib, self.lib.fq3_debug_enable(self.h, int(on)))

    def debug_layers(self, which: str, nt: int) -> Dict[str, torch.Tensor]:
        """Intermediates of the last talker step (which='t') / predictor pass 0 (which='p') as float32 CPU tensors."""
        d = self.talker_cfg if which == "t" else self.pred_cfg
        H, I, L = d["hidden_size"], d["intermediate_size"], d["num_hidden_layers"]
        qd, kd = d["num_attention_heads"] * 128, d["num_key_value_heads"] * 128
        strides = []
        for dd in (self.talker_cfg, self.pred_cfg):
            strides.append(2 * (dd["num_attention_heads"] + 2 * dd["num_key_value_heads"]) * 128 +
                           2 * dd["num_attention_heads"] * 128 + 4 * dd["hidden_size"] + 2 * dd["intermediate_size"])
        stride = max(strides)
        buf = torch.empty(stride * L, dtype=torch.float32)
        _check(self.lib, self.lib.fq3_debug_read(self.h, 0, buf.numel(), buf.data_ptr()))
        out = {}
        for l in range(L):
            r = buf[l * stride:(l + 1) * stride]
            o = 0
            out[f"L{l}.qkv"] = r[o:o + 2 * (qd + 2 * kd)].view(2, qd + 2 * kd)[:nt]; o += 2 * (qd + 2 * kd)
            out[f"L{l}.attn"] = r[o:o + nt * qd].view(nt, qd); o += 2 * qd
            out[f"L{l}.x1"] = r[o:o + 2 * H].view(2, H)[:nt]; o += 2 * H
            out[f"L{l}.act"] = r[o:o + nt * I].view(nt, I); o += 2 * I
            out[f"L{l}.x"] = r[o:o + 2 * H].view(2, H)[:nt]
        return out

    def probe_timestamps(self, n: int) -> torch.Tensor:
        buf = torch.empty(2 * n, dtype=torch.float32)
        _check(self.lib, self.lib.fq3_debug_read(self.h, 0, buf.numel(), buf.data_ptr()))
        return buf.view(torch.int64)

    def barrier_test(self, n: int, kind: int):
        _check(self.lib, self.lib.fq3_barrier_test(self.h, int(n), int(kind), self._stream()))

    def tape_bytes(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _check(self.lib, self.lib.fq3_tape_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def num_ctas(self) -> int:
        return self.lib.fq3_num_ctas(self.h)

    @property
    def launch_count(self) -> int:
        return self.lib.fq3_launch_count(self.h)
