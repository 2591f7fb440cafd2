// fq3_engine.cu -- C ABI implementation (see include/fq3_engine.h): engine lifecycle, weight-tape packing,
// launches of the persistent decode kernel (fq3_decode.cuh).  sm_100a only.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/fq3_engine.h"
#include "fq3_decode.cuh"
#include "fq3_decode_batch.cuh"

using namespace fq3;

// ------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(FQ3_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// make the engine's device current for the duration of an ABI call and restore the caller's (torch's) device after
struct DevGuard {
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};

struct SegHost {
  int rows, K;
  uint64_t bytes;  // total tape bytes of this segment (all CTAs)
};

struct PackGrp {
  uint64_t tape_off;
  uint32_t rowsrc_idx;
  uint16_t rows, m, ntiles, pad;   // fp32 layout: rows, 512-byte chunks per tile; mma layout: n_mt | kind << 8, G
  int32_t K;
};

// host-side record of one request slot (fq3_begin_request latches it; the kernels read it through KParams / SlotParams)
struct SlotHost {
  bool active = false;
  int prefill_len = 0, rope_delta = 0, n_left_pad = 0, max_new = 0, min_new = 0, trailing_len = 0;
  const void* trailing = nullptr;
  const void* tts_pad = nullptr;
  const float* uniforms = nullptr;
  Sampling sp_t{1, 50, 0.9f, 1.0f, 1.05f}, sp_p{1, 50, 0.9f, 1.0f, 1.0f};
};

struct fq3_engine {
  fq3_config cfg;
  bool bf16;
  size_t esz;
  int ncta;
  int dev;
  int max_batch = 1;
  bool loaded = false;
  std::vector<SlotHost> slots;
  size_t tkv_slot = 0, pkv_slot = 0;   // bytes of one slot's K (or V) cache: talker / predictor
  // batched decode: activation matrices + per-launch slot table
  float *XB = nullptr, *X1B = nullptr, *QKVB = nullptr, *LOGB = nullptr;
  void *XNB = nullptr, *ATTB = nullptr, *ACTB = nullptr, *PINB = nullptr;
  int* TOKB = nullptr;
  SlotParams* sl_dev = nullptr;
  SlotParams* sl_host = nullptr;  // pinned
  // device buffers (slot-major: slot s starts at s * <per-slot size>)
  void *t_kc = nullptr, *t_vc = nullptr, *p_kc = nullptr, *p_vc = nullptr;
  float *X = nullptr, *X1 = nullptr, *QKV = nullptr, *ATT = nullptr, *ACT = nullptr, *LOGITS = nullptr, *PART = nullptr;
  unsigned* bar = nullptr;
  int* state = nullptr;
  int* state_host = nullptr;  // pinned
  float* past_hidden = nullptr;
  uint32_t* seen = nullptr;
  float* dbg = nullptr;
  size_t dbg_floats = 0;
  long long dbg_stride = 0;
  int dbg_on = 0;
  uint8_t* tape = nullptr;
  size_t tape_bytes = 0;
  Grp* grps = nullptr;
  uint32_t* segtab = nullptr;
  uint32_t* cta_grp_off = nullptr;
  std::map<std::string, void*> tabs;  // owned small tables (device)
  std::vector<SegHost> segs;
  int64_t talker_step_bytes = 0, predictor_frame_bytes = 0;
  KParams kp;  // template parameters (static part)
  int64_t launches = 0;
  // K3 prefill: borrowed row-major weights + scratch
  const void *pf_qkv = nullptr, *pf_o = nullptr, *pf_gu = nullptr, *pf_down = nullptr, *pf_head = nullptr;
  void* pf_buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool pf_ready = false;
};

int g_fq3_gemm_backend = 0;  // shared with fq3_codec.cu
extern "C" int fq3_set_gemm_backend(int32_t backend) {
  g_fq3_gemm_backend = backend == 2 ? 2 : (backend ? 1 : 0);
  return 0;
}

static size_t smem_bytes() { return sizeof(Smem); }

// ------------------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__global__ void pack_kernel(const PackGrp* __restrict__ pg, int npg, const void* const* __restrict__ rowsrc,
                            uint8_t* __restrict__ tape) {
  for (int b = blockIdx.x; b < npg; b += gridDim.x) {
    const PackGrp g = pg[b];
    const int rows = g.rows, m = g.m;
    const long long total = (long long)rows * m * g.ntiles * 32;
    uint4* dst = reinterpret_cast<uint4*>(tape + g.tape_off);
    for (long long q = threadIdx.x; q < total; q += blockDim.x) {
      const int lane = (int)(q & 31);
      long long rem = q >> 5;
      const int j = (int)(rem % m);
      rem /= m;
      const int r = (int)(rem % rows);
      const int t = (int)(rem / rows);
      const int kb = t * m + j;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(rowsrc[g.rowsrc_idx + r]);
      uint4 v;
      if constexpr (BF) {
        const uint2 a = *reinterpret_cast<const uint2*>(src + ((size_t)kb * 256 + lane * 4) * 2);
        const uint2 c = *reinterpret_cast<const uint2*>(src + ((size_t)kb * 256 + 128 + lane * 4) * 2);
        v = make_uint4(a.x, a.y, c.x, c.y);
      } else {
        v = *reinterpret_cast<const uint4*>(src + ((size_t)kb * 128 + lane * 4) * 4);
      }
      dst[q] = v;
    }
  }
}

// bf16 tensor-core layout: per tile [m-tile][k-group][step 0..3][lane][16 B] holding mma.m16n8k16 A fragments
// (a0,a1,a2,a3) = rowA[kk,kk+1], rowB[kkB,kkB+1], rowA[kk+2,kk+3], rowB[kkB+2,kkB+3], kk = 64*kgroup + 16*t + 4*step.
// kind 0 FULL: rowA = r0+16*mt+g, rowB = rowA+8;  kind 1 HALF: rowA = rowB = r0+g, kkB = K/2 + kk;
// kind 2 GU: rowA = gate row, rowB = up row of pair r0/2 + 8*mt + g (rowsrc holds gate/up interleaved).
__global__ void pack_mma_kernel(const PackGrp* __restrict__ pg, int npg, const void* const* __restrict__ rowsrc,
                                uint8_t* __restrict__ tape) {
  for (int b = blockIdx.x; b < npg; b += gridDim.x) {
    const PackGrp g = pg[b];
    const int n_mt = g.rows & 0xff, kind = g.rows >> 8, G = g.m;
    const long long total = (long long)g.ntiles * n_mt * G * 128;
    uint4* dst = reinterpret_cast<uint4*>(tape + g.tape_off);
    for (long long q = threadIdx.x; q < total; q += blockDim.x) {
      const int lane = (int)(q & 31), st = (int)((q >> 5) & 3);
      long long rem = q >> 7;
      const int qq = (int)(rem % G);
      rem /= G;
      const int mt = (int)(rem % n_mt);
      const int tl = (int)(rem / n_mt);
      const int gq = lane >> 2, t = lane & 3;
      const int kk = 64 * (tl * G + qq) + 16 * t + 4 * st;
      const uint8_t *ra, *rb;
      int kb = kk;
      if (kind == 0) {
        ra = reinterpret_cast<const uint8_t*>(rowsrc[g.rowsrc_idx + mt * 16 + gq]);
        rb = reinterpret_cast<const uint8_t*>(rowsrc[g.rowsrc_idx + mt * 16 + gq + 8]);
      } else if (kind == 1) {
        ra = rb = reinterpret_cast<const uint8_t*>(rowsrc[g.rowsrc_idx + gq]);
        kb = g.K / 2 + kk;
      } else {
        ra = reinterpret_cast<const uint8_t*>(rowsrc[g.rowsrc_idx + 2 * (mt * 8 + gq)]);
        rb = reinterpret_cast<const uint8_t*>(rowsrc[g.rowsrc_idx + 2 * (mt * 8 + gq) + 1]);
      }
      const uint2 a = *reinterpret_cast<const uint2*>(ra + (size_t)kk * 2);
      const uint2 c = *reinterpret_cast<const uint2*>(rb + (size_t)kb * 2);
      dst[q] = make_uint4(a.x, c.x, a.y, c.y);
    }
  }
}

template <bool BF>
__global__ void set_state_kernel(int* state, float* past_hidden, uint32_t* seen, const void* ph_src, int Ht,
                                 int token, int gen_step) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) {
    state[0] = token; state[1] = 0; state[2] = gen_step; state[3] = 0; state[4] = 0;
  }
  for (int k = tid; k < Ht; k += gridDim.x * blockDim.x) past_hidden[k] = ldw<BF>(ph_src, k);
  for (int k = tid; k < VMAX / 32; k += gridDim.x * blockDim.x) seen[k] = 0u;
}

template <bool BF>
__global__ void get_hidden_kernel(const float* past_hidden, void* dst, int Ht) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < Ht; k += gridDim.x * blockDim.x) stw<BF>(dst, k, past_hidden[k]);
}

// standalone sampler: one CTA of 256 threads running the same sample_block the fused loop uses
template <bool BF>
__global__ void __launch_bounds__(NCT, 1)
    sample_kernel(const __grid_constant__ KParams P, const void* logits, int V, Sampling sp, float u,
                  const long long* hist, int n_hist, int suppress_special, int eos, int suppress_eos,
                  long long* out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < VMAX / 32; i += NCT) s.seen[i] = 0u;
  __syncthreads();
  for (int i = tid; i < n_hist; i += NCT) {
    const int v = (int)hist[i];
    if (v >= 0 && v < V) atomicOr(&s.seen[v >> 5], 1u << (v & 31));
  }
  for (int v = tid; v < V; v += NCT) P.LOGITS[v] = ldw<BF>(logits, v);
  __threadfence();
  __syncthreads();
  Ctx c{P, tid, tid >> 5, tid & 31, 0u, 0u};
  SampleArgs a;
  a.logits = P.LOGITS; a.V = V; a.sp = sp; a.u = u;
  a.use_penalty = true;
  a.sup0 = suppress_special ? (V > 1024 ? V - 1024 : 0) : V;
  a.suppress_eos = suppress_eos != 0;
  a.eos = eos;
  const int tok = sample_block<BF>(c, a);
  if (tid == 0) out[0] = tok;
}

// ------------------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------------------
static int check_stack(const fq3_stack_config& s, const char* nm, int nt) {
  const int qd = s.num_attention_heads * 128, kd = s.num_key_value_heads * 128;
  if (s.hidden_size <= 0 || s.hidden_size % 256 || s.intermediate_size % 256 || qd % 256 || kd % 128)
    return fail(FQ3_ERR_INVALID, "%s: hidden/intermediate/q dims must be multiples of 256 (head_dim fixed at 128)", nm);
  if (s.num_attention_heads % s.num_key_value_heads) return fail(FQ3_ERR_INVALID, "%s: heads %% kv_heads != 0", nm);
  const int mx = std::max(std::max(s.hidden_size, qd), s.intermediate_size);
  if (nt * mx > XS_FLOATS) return fail(FQ3_ERR_INVALID, "%s: %d x max(H,qd,I)=%d exceeds the %d-float staging buffer", nm, nt, mx, XS_FLOATS);
  if (s.vocab_size > VMAX || s.vocab_size % 2) return fail(FQ3_ERR_INVALID, "%s: vocab_size must be even and <= %d", nm, VMAX);
  if (s.num_hidden_layers <= 0) return fail(FQ3_ERR_INVALID, "%s: num_hidden_layers", nm);
  return 0;
}

extern "C" int fq3_engine_create(const fq3_config* cfg, fq3_engine** out) {
  if (!cfg || !out) return fail(FQ3_ERR_INVALID, "null argument");
  if (cfg->dtype != FQ3_F32 && cfg->dtype != FQ3_BF16) return fail(FQ3_ERR_INVALID, "dtype must be FQ3_F32 or FQ3_BF16");
  int rc;
  if ((rc = check_stack(cfg->talker, "talker", 1))) return rc;
  if ((rc = check_stack(cfg->predictor, "predictor", 2))) return rc;
  if (cfg->talker.hidden_size > HMAX) return fail(FQ3_ERR_INVALID, "talker hidden_size > %d", HMAX);
  if (2 * cfg->talker.hidden_size > XS_FLOATS) return fail(FQ3_ERR_INVALID, "talker hidden too large for mtp staging");
  if (cfg->max_seq_len < 8 || cfg->max_seq_len > SEQMAX) return fail(FQ3_ERR_INVALID, "max_seq_len must be in [8,%d]", SEQMAX);
  if (cfg->num_code_groups < 2 || cfg->num_code_groups > 16) return fail(FQ3_ERR_INVALID, "num_code_groups must be in [2,16]");
  if (cfg->rope_positions < cfg->max_seq_len) return fail(FQ3_ERR_INVALID, "rope_positions < max_seq_len");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(FQ3_ERR_INVALID, "device %d out of range", cfg->device);
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) return fail(FQ3_ERR_INVALID, "sm_100a required (device is sm_%d%d)", prop.major, prop.minor);
  int coop = 0;
  CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device));
  if (!coop) return fail(FQ3_ERR_INVALID, "device lacks cooperative launch");
  fq3_engine* e = new fq3_engine();
  e->cfg = *cfg;
  e->bf16 = cfg->dtype == FQ3_BF16;
  e->esz = e->bf16 ? 2 : 4;
  e->dev = cfg->device;
  e->ncta = cfg->num_ctas > 0 ? std::min(cfg->num_ctas, prop.multiProcessorCount) : prop.multiProcessorCount;
  e->max_batch = cfg->max_batch > 0 ? cfg->max_batch : 1;
  if (e->max_batch > MAXB) { delete e; return fail(FQ3_ERR_INVALID, "max_batch %d exceeds %d", cfg->max_batch, MAXB); }
  if (e->bf16) {
    CK(cudaFuncSetAttribute(fq3_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    CK(cudaFuncSetAttribute(fq3_decode_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    CK(cudaFuncSetAttribute(sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  } else {
    CK(cudaFuncSetAttribute(fq3_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    CK(cudaFuncSetAttribute(fq3_decode_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
    CK(cudaFuncSetAttribute(sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  }
  int occ = 0;
  if (e->bf16) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fq3_decode_kernel<true>, NTHREADS, smem_bytes()));
  else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fq3_decode_kernel<false>, NTHREADS, smem_bytes()));
  if (occ < 1) { delete e; return fail(FQ3_ERR_INVALID, "decode kernel does not fit on an SM (smem %zu)", smem_bytes()); }

  const fq3_stack_config &T = cfg->talker, &Pc = cfg->predictor;
  const int MB = e->max_batch;
  e->slots.assign(MB, SlotHost());
  const size_t tkv = (size_t)T.num_hidden_layers * T.num_key_value_heads * cfg->max_seq_len * 128 * e->esz;
  const size_t pkv = (size_t)Pc.num_hidden_layers * Pc.num_key_value_heads * 32 * 128 * e->esz;
  e->tkv_slot = tkv; e->pkv_slot = pkv;
  CK(cudaMalloc(&e->t_kc, tkv * MB)); CK(cudaMalloc(&e->t_vc, tkv * MB));
  CK(cudaMalloc(&e->p_kc, pkv * MB)); CK(cudaMalloc(&e->p_vc, pkv * MB));
  CK(cudaMemset(e->t_kc, 0, tkv * MB)); CK(cudaMemset(e->t_vc, 0, tkv * MB));
  CK(cudaMemset(e->p_kc, 0, pkv * MB)); CK(cudaMemset(e->p_vc, 0, pkv * MB));
  const int ldX = std::max(T.hidden_size, Pc.hidden_size);
  const int ldQKV = std::max((T.num_attention_heads + 2 * T.num_key_value_heads) * 128,
                             (Pc.num_attention_heads + 2 * Pc.num_key_value_heads) * 128);
  const int ldATT = std::max(T.num_attention_heads, Pc.num_attention_heads) * 128;
  const int ldACT = std::max(T.intermediate_size, Pc.intermediate_size);
  CK(cudaMalloc(&e->X, 2 * ldX * sizeof(float))); CK(cudaMalloc(&e->X1, 2 * ldX * sizeof(float)));
  CK(cudaMalloc(&e->QKV, 2 * ldQKV * sizeof(float))); CK(cudaMalloc(&e->ATT, 2 * ldATT * sizeof(float)));
  CK(cudaMalloc(&e->ACT, 2 * ldACT * sizeof(float))); CK(cudaMalloc(&e->LOGITS, VMAX * sizeof(float)));
  CK(cudaMalloc(&e->bar, 32768)); CK(cudaMemset(e->bar, 0, 32768));
  CK(cudaMalloc(&e->PART, (size_t)T.num_attention_heads * 16 * PART_STRIDE * sizeof(float)));
  CK(cudaMalloc(&e->state, 32 * MB)); CK(cudaMemset(e->state, 0, 32 * MB));   // 8 ints per slot
  CK(cudaMallocHost(&e->state_host, 32 * MB));
  CK(cudaMalloc(&e->past_hidden, (size_t)MB * HMAX * sizeof(float))); CK(cudaMemset(e->past_hidden, 0, (size_t)MB * HMAX * sizeof(float)));
  CK(cudaMalloc(&e->seen, (size_t)MB * (VMAX / 8))); CK(cudaMemset(e->seen, 0, (size_t)MB * (VMAX / 8)));
  if (MB > 1) {
    // batched decode: activation matrices [column][ld] (column = slot, predictor pass 0: token * B + slot)
    auto zalloc = [&](void** p, size_t bytes) -> cudaError_t {
      cudaError_t r = cudaMalloc(p, bytes);
      return r != cudaSuccess ? r : cudaMemset(*p, 0, bytes);
    };
    CK(zalloc((void**)&e->XB, (size_t)MAXCOL * ldX * sizeof(float)));
    CK(zalloc((void**)&e->X1B, (size_t)MAXCOL * ldX * sizeof(float)));
    CK(zalloc((void**)&e->QKVB, (size_t)MAXCOL * ldQKV * sizeof(float)));
    CK(zalloc((void**)&e->LOGB, (size_t)MAXB * VMAX * sizeof(float)));
    CK(zalloc(&e->XNB, (size_t)MAXCOL * ldX * e->esz));
    CK(zalloc(&e->ATTB, (size_t)MAXCOL * ldATT * e->esz));
    CK(zalloc(&e->ACTB, (size_t)MAXCOL * ldACT * e->esz));
    CK(zalloc(&e->PINB, (size_t)MAXCOL * HMAX * e->esz));
    CK(zalloc((void**)&e->TOKB, MAXB * sizeof(int)));
    CK(cudaMalloc(&e->sl_dev, MAXB * sizeof(SlotParams)));
    CK(cudaMallocHost(&e->sl_host, MAXB * sizeof(SlotParams)));
  }
  {
    const long long rec_t = 2LL * (T.num_attention_heads + 2 * T.num_key_value_heads) * 128 + 2LL * T.num_attention_heads * 128 + 4LL * T.hidden_size + 2LL * T.intermediate_size;
    const long long rec_p = 2LL * (Pc.num_attention_heads + 2 * Pc.num_key_value_heads) * 128 + 2LL * Pc.num_attention_heads * 128 + 4LL * Pc.hidden_size + 2LL * Pc.intermediate_size;
    e->dbg_stride = std::max(rec_t, rec_p);
    e->dbg_floats = (size_t)e->dbg_stride * std::max(T.num_hidden_layers, Pc.num_hidden_layers);
    CK(cudaMalloc(&e->dbg, e->dbg_floats * sizeof(float)));
    CK(cudaMemset(e->dbg, 0, e->dbg_floats * sizeof(float)));
  }
  KParams& k = e->kp;
  memset(&k, 0, sizeof(k));
  auto fill = [&](StackDev& s, const fq3_stack_config& c, void* kc, void* vc, int S) {
    s.H = c.hidden_size; s.I = c.intermediate_size; s.L = c.num_hidden_layers;
    s.nH = c.num_attention_heads; s.nKV = c.num_key_value_heads; s.V = c.vocab_size;
    s.qd = s.nH * 128; s.kd = s.nKV * 128; s.rep = s.nH / s.nKV; s.eps = c.rms_norm_eps;
    s.kc = kc; s.vc = vc; s.S = S;
  };
  fill(k.t, T, e->t_kc, e->t_vc, cfg->max_seq_len);
  fill(k.p, Pc, e->p_kc, e->p_vc, 32);
  k.ncta = e->ncta;
  k.X = e->X; k.X1 = e->X1; k.QKV = e->QKV; k.ATT = e->ATT; k.ACT = e->ACT; k.LOGITS = e->LOGITS;
  k.ldX = ldX; k.ldQKV = ldQKV; k.ldATT = ldATT; k.ldACT = ldACT;
  k.bar = e->bar; k.state = e->state; k.past_hidden = e->past_hidden; k.seen = e->seen;
  k.XB = e->XB; k.X1B = e->X1B; k.QKVB = e->QKVB; k.LOGB = e->LOGB;
  k.XNB = e->XNB; k.ATTB = e->ATTB; k.ACTB = e->ACTB; k.PINB = e->PINB; k.TOKB = e->TOKB;
  k.nslots = 0; k.sl = e->sl_dev;
  k.has_mtp = cfg->has_mtp_projection; k.ncb = cfg->num_code_groups - 1; k.eos = cfg->codec_eos_token_id;
  k.max_seq_len = cfg->max_seq_len;
  k.dbg = e->dbg; k.dbg_stride_layer = e->dbg_stride;
  k.pred_pin_layers = 2;
  if (const char* v = getenv("FQ3_PRED_PIN")) k.pred_pin_layers = std::max(atoi(v), 0);   // tuning knob
  {
    // split-key talker attention (bf16 engines): S CTAs per q-head; a slice must fit the 4 ring tiles it may hold
    int Sx = e->bf16 ? std::min(e->ncta / std::max(T.num_attention_heads, 1), 16) : 0;
    if (Sx < 2 || (cfg->max_seq_len + Sx - 1) / Sx > 4 * KVT_KEYS) Sx = 0;
    if (const char* v = getenv("FQ3_ATTN_SPLIT")) Sx = std::min(Sx, std::max(atoi(v), 0)) < 2 ? 0 : std::min(Sx, atoi(v));
    k.attn_split = Sx;
    k.attn_split_min = 192;   // measured on B200 (1.7B geometry): split wins from ~250 cached keys up (0.93 vs 1.01 ms/step at 300)
    if (const char* v = getenv("FQ3_ATTN_SPLIT_MIN")) k.attn_split_min = std::max(atoi(v), 0);
    k.PART = e->PART;
    k.attn_cnt = e->bar + 1024;
  }
  k.sp_t = Sampling{1, 50, 0.9f, 1.0f, 1.05f};
  k.sp_p = Sampling{1, 50, 0.9f, 1.0f, 1.0f};
  *out = e;
  return 0;
}

extern "C" void fq3_engine_destroy(fq3_engine* e) {
  if (!e) return;
  cudaSetDevice(e->dev);
  void* ptrs[] = {e->t_kc, e->t_vc, e->p_kc, e->p_vc, e->X, e->X1, e->QKV, e->ATT, e->ACT, e->LOGITS, e->PART, e->bar,
                  e->state, e->past_hidden, e->seen, e->dbg, e->tape, e->grps, e->segtab, e->cta_grp_off,
                  e->XB, e->X1B, e->QKVB, e->LOGB, e->XNB, e->ATTB, e->ACTB, e->PINB, e->TOKB, e->sl_dev};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& kv : e->tabs)
    if (kv.second) cudaFree(kv.second);
  for (void* p : e->pf_buf)
    if (p) cudaFree(p);
  if (e->state_host) cudaFreeHost(e->state_host);
  if (e->sl_host) cudaFreeHost(e->sl_host);
  delete e;
}

// table[r][o] = round(bias[o] + sum_k W[o][k] * emb[r][k]) for every row r of the 15 predictor codec embeddings:
// the code predictor's input projection (predictor_graph.py:53 small_to_mtp_projection) of an embedding row depends
// only on the code, so it is tabulated once per weight load instead of being recomputed (GEMV + grid barrier) in 14
// of the 15 passes of every frame.  4 embedding rows per block, one warp per output row.
template <bool BF>
__global__ void mtp_table_kernel(const void* __restrict__ emb, const void* __restrict__ W, const void* __restrict__ bias,
                                 void* __restrict__ table, int K, int N, long long rows) {
  extern __shared__ float es[];  // [4][K]
  const long long r0 = (long long)blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) {
    const long long r = r0 + i / K;
    es[i] = r < rows ? ldw<BF>(emb, (size_t)r * K + (i % K)) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < N; o += nw) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K; k += 32) {
      const float w = ldw<BF>(W, (size_t)o * K + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(w, es[j * K + k], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int s = 16; s; s >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], s);
    if (lane < 4 && r0 + lane < rows) {
      const float b = bias ? ldw<BF>(bias, o) : 0.f;
      const float v = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
      stw<BF>(table, (size_t)(r0 + lane) * N + o, rnd<BF>(v + b));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// weights: copy tables, build segments and the per-CTA tape
// ------------------------------------------------------------------------------------------------------------
extern "C" int fq3_engine_load_weights(fq3_engine* e, const fq3_tensor* tensors, int32_t n, void* stream_) {
  if (!e || !tensors) return fail(FQ3_ERR_INVALID, "null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  DevGuard dev_guard(e->dev);
  std::map<std::string, const fq3_tensor*> tm;
  for (int i = 0; i < n; ++i) tm[tensors[i].name] = &tensors[i];
  const fq3_config& cfg = e->cfg;
  const size_t esz = e->esz;
  auto need = [&](const std::string& nm, int64_t numel, const void** ptr) -> int {
    auto it = tm.find(nm);
    if (it == tm.end()) return fail(FQ3_ERR_INVALID, "missing tensor '%s'", nm.c_str());
    if (it->second->numel != numel)
      return fail(FQ3_ERR_INVALID, "tensor '%s': numel %lld, expected %lld", nm.c_str(), (long long)it->second->numel, (long long)numel);
    *ptr = it->second->dev_ptr;
    return 0;
  };
  int rc;
  // ---- small tables: copied into engine-owned storage
  auto own = [&](const std::string& nm, int64_t numel, size_t elem, const void** devp) -> int {
    const void* src;
    if ((rc = need(nm, numel, &src))) return rc;
    void* dst = nullptr;
    auto it = e->tabs.find(nm);
    if (it != e->tabs.end() && it->second) cudaFree(it->second);
    CK(cudaMalloc(&dst, (size_t)numel * elem));
    CK(cudaMemcpyAsync(dst, src, (size_t)numel * elem, cudaMemcpyDeviceToDevice, stream));
    e->tabs[nm] = dst;
    *devp = dst;
    return 0;
  };
  KParams& k = e->kp;
  struct StackNames { const char* pre; StackDev* s; const fq3_stack_config* c; };
  StackNames sn[2] = {{"t.", &k.t, &cfg.talker}, {"p.", &k.p, &cfg.predictor}};
  for (auto& s : sn) {
    const std::string p = s.pre;
    const int L = s.c->num_hidden_layers, H = s.c->hidden_size;
    if ((rc = own(p + "ln_in", (int64_t)L * H, esz, &s.s->ln_in))) return rc;
    if ((rc = own(p + "ln_post", (int64_t)L * H, esz, &s.s->ln_post))) return rc;
    if ((rc = own(p + "qnorm", (int64_t)L * 128, esz, &s.s->qnorm))) return rc;
    if ((rc = own(p + "knorm", (int64_t)L * 128, esz, &s.s->knorm))) return rc;
    if ((rc = own(p + "ln_f", H, esz, &s.s->ln_f))) return rc;
  }
  {
    const void* p;
    if ((rc = own("t.cos", (int64_t)cfg.rope_positions * 128, 4, &p))) return rc; k.t.cos = (const float*)p;
    if ((rc = own("t.sin", (int64_t)cfg.rope_positions * 128, 4, &p))) return rc; k.t.sin = (const float*)p;
    k.t.npos = cfg.rope_positions;
    if ((rc = own("p.cos", 32 * 128, 4, &p))) return rc; k.p.cos = (const float*)p;
    if ((rc = own("p.sin", 32 * 128, 4, &p))) return rc; k.p.sin = (const float*)p;
    k.p.npos = 32;
    const int Ht = cfg.talker.hidden_size;
    if ((rc = own("t.embed", (int64_t)cfg.talker.vocab_size * Ht, esz, &k.t_embed))) return rc;
    if ((rc = own("p.embeds", (int64_t)k.ncb * cfg.predictor.vocab_size * Ht, esz, &k.p_embeds))) return rc;
    k.mtp_tab = nullptr;
    if (cfg.has_mtp_projection) {
      if ((rc = own("p.mtp_b", cfg.predictor.hidden_size, esz, &k.mtp_b))) return rc;
      const char* off = getenv("FQ3_NO_MTP_TABLE");
      const size_t sm = (size_t)4 * Ht * sizeof(float);
      if (!(off && off[0] == '1') && sm <= 48 * 1024) {
        const void* wm;
        const int Hp = cfg.predictor.hidden_size;
        if ((rc = need("p.mtp_w", (int64_t)Hp * Ht, &wm))) return rc;
        const long long rows = (long long)k.ncb * cfg.predictor.vocab_size;
        void* tab = nullptr;
        auto it = e->tabs.find("p.mtp_table");
        if (it != e->tabs.end() && it->second) cudaFree(it->second);
        CK(cudaMalloc(&tab, (size_t)rows * Hp * esz));
        e->tabs["p.mtp_table"] = tab;
        const unsigned nb = (unsigned)((rows + 3) / 4);
        if (e->bf16) mtp_table_kernel<true><<<nb, 256, sm, stream>>>(k.p_embeds, wm, k.mtp_b, tab, Ht, Hp, rows);
        else mtp_table_kernel<false><<<nb, 256, sm, stream>>>(k.p_embeds, wm, k.mtp_b, tab, Ht, Hp, rows);
        CK(cudaGetLastError());
        k.mtp_tab = tab;
      }
    } else {
      k.mtp_b = nullptr;
    }
  }
  // ---- segments and their row sources
  std::vector<SegHost>& segs = e->segs;
  segs.clear();
  std::vector<const void*> rowsrc;          // concatenated row pointers
  std::vector<uint32_t> seg_rowsrc0;        // first index into rowsrc per segment
  auto add_seg = [&](int rows, int K) {
    segs.push_back(SegHost{rows, K, 0});
    seg_rowsrc0.push_back((uint32_t)rowsrc.size());
  };
  auto push_rows = [&](const void* base, int64_t row0, int rows, int K, int stride_rows = 1, int start = 0) {
    (void)stride_rows; (void)start;
    for (int r = 0; r < rows; ++r) rowsrc.push_back((const uint8_t*)base + (size_t)(row0 + r) * K * esz);
  };
  int seg_id = 0;
  for (int si = 0; si < 2; ++si) {
    const std::string p = sn[si].pre;
    const fq3_stack_config& c = *sn[si].c;
    const int L = c.num_hidden_layers, H = c.hidden_size, I = c.intermediate_size;
    const int qd = c.num_attention_heads * 128, kd = c.num_key_value_heads * 128;
    const void *wq, *wk, *wv, *wo, *wg, *wu, *wd;
    if ((rc = need(p + "q", (int64_t)L * qd * H, &wq))) return rc;
    if ((rc = need(p + "k", (int64_t)L * kd * H, &wk))) return rc;
    if ((rc = need(p + "v", (int64_t)L * kd * H, &wv))) return rc;
    if ((rc = need(p + "o", (int64_t)L * H * qd, &wo))) return rc;
    if ((rc = need(p + "gate", (int64_t)L * I * H, &wg))) return rc;
    if ((rc = need(p + "up", (int64_t)L * I * H, &wu))) return rc;
    if ((rc = need(p + "down", (int64_t)L * H * I, &wd))) return rc;
    sn[si].s->seg_base = seg_id;
    for (int l = 0; l < L; ++l) {
      add_seg(qd + 2 * kd, H);
      push_rows(wq, (int64_t)l * qd, qd, H);
      push_rows(wk, (int64_t)l * kd, kd, H);
      push_rows(wv, (int64_t)l * kd, kd, H);
      add_seg(H, qd);
      push_rows(wo, (int64_t)l * H, H, qd);
      add_seg(2 * I, H);
      for (int r = 0; r < I; ++r) {
        rowsrc.push_back((const uint8_t*)wg + ((size_t)l * I + r) * H * esz);
        rowsrc.push_back((const uint8_t*)wu + ((size_t)l * I + r) * H * esz);
      }
      add_seg(H, I);
      push_rows(wd, (int64_t)l * H, H, I);
      seg_id += 4;
    }
    if (si == 0) {
      const void* wh;
      if ((rc = need("t.head", (int64_t)c.vocab_size * H, &wh))) return rc;
      sn[si].s->seg_head = seg_id;
      add_seg(c.vocab_size, H);
      push_rows(wh, 0, c.vocab_size, H);
      seg_id += 1;
    } else {
      const void* wh;
      if ((rc = need("p.heads", (int64_t)k.ncb * c.vocab_size * H, &wh))) return rc;
      sn[si].s->seg_head = seg_id;
      for (int i = 0; i < k.ncb; ++i) {
        add_seg(c.vocab_size, H);
        push_rows(wh, (int64_t)i * c.vocab_size, c.vocab_size, H);
        seg_id += 1;
      }
      if (cfg.has_mtp_projection) {
        const void* wm;
        const int Ht = cfg.talker.hidden_size;
        if ((rc = need("p.mtp_w", (int64_t)H * Ht, &wm))) return rc;
        k.seg_mtp = seg_id;
        add_seg(H, Ht);
        push_rows(wm, 0, H, Ht);
        seg_id += 1;
      } else {
        k.seg_mtp = -1;
      }
    }
  }
  const int nseg = (int)segs.size();
  if (nseg > MAXSEG) return fail(FQ3_ERR_INVALID, "too many segments (%d > %d)", nseg, MAXSEG);
  k.nseg = nseg;
  // ---- distribute row pairs over CTAs, split into groups, lay out the tape
  const int ncta = e->ncta;
  const int EPW = e->bf16 ? 256 : 128;  // elements per 512-byte warp read
  std::vector<std::vector<Grp>> cta_grps(ncta);
  std::vector<uint32_t> segtab((size_t)ncta * nseg, 0);
  std::vector<PackGrp> pack;
  // per-CTA byte totals to place groups: first pass collects sizes
  struct Tmp { int cta, seg, row0, rows, m, ntiles; uint64_t bytes; uint32_t rsrc; };
  std::vector<Tmp> tmp;
  int rot = 0;
  // which segments are gate/up (interleaved) segments
  std::vector<char> seg_is_gu(nseg, 0);
  for (int si = 0; si < 2; ++si)
    for (int l = 0; l < sn[si].c->num_hidden_layers; ++l) seg_is_gu[sn[si].s->seg_base + 4 * l + 2] = 1;
  for (int sg = 0; sg < nseg; ++sg) {
    const int rows = segs[sg].rows, K = segs[sg].K;
    segs[sg].bytes = (uint64_t)rows * K * esz;
    if (e->bf16) {
      // ---- tensor-core layout: units of 8 rows (plain) or 8 gate/up pairs (GU)
      const bool gu = seg_is_gu[sg];
      const int unit_rows = gu ? 16 : 8;
      if (rows % unit_rows || K % 128) return fail(FQ3_ERR_INVALID, "segment %d: rows %d / K %d not tileable for the bf16 tensor-core tape", sg, rows, K);
      const int units = rows / unit_rows, base = units / ncta, extra = units % ncta;
      int unit0 = 0;
      for (int c = 0; c < ncta; ++c) {
        const int uc = base + ((((c - rot) % ncta + ncta) % ncta) < extra ? 1 : 0);
        const int begin = (int)cta_grps[c].size();
        int ng = 0;
        auto emit = [&](int kind, int n_mt, int row0, uint32_t rsrc) {
          const int Keff = kind == 1 ? K / 2 : K;
          const int KG = Keff / 64;
          int G = 1;
          for (int d = 1; d <= KG; ++d)
            if (KG % d == 0 && n_mt * d <= 16) G = d;
          Tmp t{c, sg, row0, n_mt | (kind << 8), G, KG / G, (uint64_t)n_mt * KG * 2048, rsrc};
          tmp.push_back(t);
          Grp g; g.off16 = 0; g.row0 = row0; g.rows = (uint16_t)(n_mt | (kind << 8)); g.m = (uint16_t)G;
          g.ntiles = (uint16_t)(KG / G); g.pad = 0;
          cta_grps[c].push_back(g);
          ++ng;
        };
        if (gu) {
          for (int u = 0; u < uc; u += 2) {
            const int n_mt = std::min(2, uc - u);
            const int pair0 = (unit0 + u) * 8;
            emit(2, n_mt, pair0, seg_rowsrc0[sg] + 2u * pair0);
          }
        } else {
          const int nfull = uc / 2;
          for (int f = 0; f < nfull; f += 2) {
            const int n_mt = std::min(2, nfull - f);
            const int r0 = (unit0 + 2 * f) * 8;
            emit(0, n_mt, r0, seg_rowsrc0[sg] + (uint32_t)r0);
          }
          if (uc % 2) {
            const int r0 = (unit0 + uc - 1) * 8;
            emit(1, 1, r0, seg_rowsrc0[sg] + (uint32_t)r0);
          }
        }
        if (ng > 255) return fail(FQ3_ERR_INVALID, "segment %d: too many groups per CTA", sg);
        segtab[(size_t)c * nseg + sg] = ((uint32_t)begin << 8) | (uint32_t)ng;
        unit0 += uc;
      }
      rot = (rot + extra) % ncta;
      continue;
    }
    if (rows % 2 || K % EPW) return fail(FQ3_ERR_INVALID, "segment %d: rows %d must be even and K %d a multiple of %d", sg, rows, K, EPW);
    const int KB = K / EPW;
    const int pairs = rows / 2, base = pairs / ncta, extra = pairs % ncta;
    int row = 0;
    for (int c = 0; c < ncta; ++c) {
      const int pc = base + ((((c - rot) % ncta + ncta) % ncta) < extra ? 1 : 0);
      const int rows_c = 2 * pc;
      const int ng = (rows_c + 31) / 32;
      int begin = (int)cta_grps[c].size();
      int r0 = row;
      for (int gi = 0; gi < ng; ++gi) {
        const int gp = pc / ng + (gi < pc % ng ? 1 : 0);
        const int gr = 2 * gp;
        int m = 1;
        for (int d = 1; d <= KB; ++d)
          if (KB % d == 0 && (size_t)gr * 512 * d <= (size_t)STAGE_BYTES) m = d;
        const int ntiles = KB / m;
        Tmp t{c, sg, r0, gr, m, ntiles, (uint64_t)gr * K * esz, seg_rowsrc0[sg] + (uint32_t)r0};
        tmp.push_back(t);
        Grp g; g.off16 = 0; g.row0 = r0; g.rows = (uint16_t)gr; g.m = (uint16_t)m; g.ntiles = (uint16_t)ntiles; g.pad = 0;
        cta_grps[c].push_back(g);
        r0 += gr;
      }
      if (ng > 255) return fail(FQ3_ERR_INVALID, "segment %d: too many groups per CTA", sg);
      segtab[(size_t)c * nseg + sg] = ((uint32_t)begin << 8) | (uint32_t)ng;
      row += rows_c;
    }
    rot = (rot + extra) % ncta;
  }
  // tape offsets: CTA-major, segment order
  std::vector<uint64_t> cta_base(ncta + 1, 0);
  {
    std::vector<uint64_t> sz(ncta, 0);
    for (auto& t : tmp) sz[t.cta] += t.bytes;
    for (int c = 0; c < ncta; ++c) cta_base[c + 1] = cta_base[c] + ((sz[c] + 1023) / 1024) * 1024;
  }
  const uint64_t tape_bytes = cta_base[ncta];
  if (tape_bytes / 16 > 0xffffffffull) return fail(FQ3_ERR_INVALID, "tape too large");
  {
    std::vector<uint64_t> cur(cta_base.begin(), cta_base.end() - 1);
    std::vector<int> gidx(ncta, 0);
    for (auto& t : tmp) {
      Grp& g = cta_grps[t.cta][gidx[t.cta]++];
      g.off16 = (uint32_t)(cur[t.cta] / 16);
      PackGrp pg; pg.tape_off = cur[t.cta]; pg.rowsrc_idx = t.rsrc;
      pg.rows = (uint16_t)t.rows; pg.m = (uint16_t)t.m; pg.ntiles = (uint16_t)t.ntiles; pg.pad = 0; pg.K = segs[t.seg].K;
      pack.push_back(pg);
      cur[t.cta] += t.bytes;
    }
  }
  std::vector<uint32_t> goff(ncta + 1, 0);
  std::vector<Grp> allg;
  for (int c = 0; c < ncta; ++c) {
    if ((int)cta_grps[c].size() > MAXGRP) return fail(FQ3_ERR_INVALID, "CTA %d has %zu row groups (> %d)", c, cta_grps[c].size(), MAXGRP);
    goff[c + 1] = goff[c] + (uint32_t)cta_grps[c].size();
    allg.insert(allg.end(), cta_grps[c].begin(), cta_grps[c].end());
  }
  // ---- upload tables, pack
  if (e->tape) { cudaFree(e->tape); e->tape = nullptr; }
  if (e->grps) { cudaFree(e->grps); e->grps = nullptr; }
  if (e->segtab) { cudaFree(e->segtab); e->segtab = nullptr; }
  if (e->cta_grp_off) { cudaFree(e->cta_grp_off); e->cta_grp_off = nullptr; }
  CK(cudaMalloc(&e->tape, tape_bytes));
  e->tape_bytes = tape_bytes;
  CK(cudaMalloc(&e->grps, allg.size() * sizeof(Grp)));
  CK(cudaMalloc(&e->segtab, segtab.size() * sizeof(uint32_t)));
  CK(cudaMalloc(&e->cta_grp_off, goff.size() * sizeof(uint32_t)));
  CK(cudaMemcpyAsync(e->grps, allg.data(), allg.size() * sizeof(Grp), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(e->segtab, segtab.data(), segtab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(e->cta_grp_off, goff.data(), goff.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  void** d_rowsrc = nullptr;
  PackGrp* d_pack = nullptr;
  CK(cudaMalloc(&d_rowsrc, rowsrc.size() * sizeof(void*)));
  CK(cudaMalloc(&d_pack, pack.size() * sizeof(PackGrp)));
  CK(cudaMemcpyAsync(d_rowsrc, rowsrc.data(), rowsrc.size() * sizeof(void*), cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_pack, pack.data(), pack.size() * sizeof(PackGrp), cudaMemcpyHostToDevice, stream));
  {
    const int grid = (int)std::min<size_t>(pack.size(), 148 * 16);
    if (e->bf16) pack_mma_kernel<<<grid, 256, 0, stream>>>(d_pack, (int)pack.size(), (const void* const*)d_rowsrc, e->tape);
    else pack_kernel<false><<<grid, 256, 0, stream>>>(d_pack, (int)pack.size(), (const void* const*)d_rowsrc, e->tape);
    e->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(stream));
  cudaFree(d_rowsrc);
  cudaFree(d_pack);
  k.mma_tape = e->bf16 ? 1 : 0;
  k.tape = e->tape; k.grps = e->grps; k.segtab = e->segtab; k.cta_grp_off = e->cta_grp_off;
  // ---- byte accounting (algorithmic bytes)
  {
    uint64_t tl = 0, pl = 0, ph = 0, pm = 0;
    for (int l = 0; l < k.t.L * 4; ++l) tl += segs[k.t.seg_base + l].bytes;
    tl += segs[k.t.seg_head].bytes;
    for (int l = 0; l < k.p.L * 4; ++l) pl += segs[k.p.seg_base + l].bytes;
    for (int i = 0; i < k.ncb; ++i) ph += segs[k.p.seg_head + i].bytes;
    if (k.seg_mtp >= 0) pm = segs[k.seg_mtp].bytes;
    e->talker_step_bytes = (int64_t)tl;
    e->predictor_frame_bytes = (int64_t)(k.ncb * pl + (k.mtp_tab ? 1 : k.ncb) * pm + ph);
  }
  e->loaded = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------------------
static int launch_decode(fq3_engine* e, const KParams& kp, cudaStream_t stream) {
  CK(cudaMemsetAsync(e->bar, 0, 32768, stream));
  void* args[] = {(void*)&kp};
  if (e->bf16)
    CK(cudaLaunchCooperativeKernel((const void*)fq3_decode_kernel<true>, dim3(e->ncta), dim3(NTHREADS), args, smem_bytes(), stream));
  else
    CK(cudaLaunchCooperativeKernel((const void*)fq3_decode_kernel<false>, dim3(e->ncta), dim3(NTHREADS), args, smem_bytes(), stream));
  e->launches++;
  return 0;
}

static Sampling to_sampling(const fq3_sampling* s) {
  Sampling r;
  r.do_sample = s->do_sample; r.top_k = s->top_k; r.temperature = s->temperature; r.top_p = s->top_p;
  r.penalty = s->repetition_penalty;
  return r;
}

static int check_slot(fq3_engine* e, int slot) {
  if (slot < 0 || slot >= e->max_batch) return fail(FQ3_ERR_INVALID, "slot %d outside [0, max_batch=%d)", slot, e->max_batch);
  return 0;
}
static void* slot_tk(fq3_engine* e, int s) { return (uint8_t*)e->t_kc + (size_t)s * e->tkv_slot; }
static void* slot_tv(fq3_engine* e, int s) { return (uint8_t*)e->t_vc + (size_t)s * e->tkv_slot; }
static void* slot_pk(fq3_engine* e, int s) { return (uint8_t*)e->p_kc + (size_t)s * e->pkv_slot; }
static void* slot_pv(fq3_engine* e, int s) { return (uint8_t*)e->p_vc + (size_t)s * e->pkv_slot; }

// kernel parameters of a single-sequence launch on slot s: the slot's caches / state + the request it latched
static KParams kp_for_slot(fq3_engine* e, int s) {
  KParams kp = e->kp;
  const SlotHost& h = e->slots[s];
  kp.t.kc = slot_tk(e, s); kp.t.vc = slot_tv(e, s); kp.p.kc = slot_pk(e, s); kp.p.vc = slot_pv(e, s);
  kp.state = e->state + 8 * s;
  kp.past_hidden = e->past_hidden + (size_t)s * HMAX;
  kp.seen = e->seen + (size_t)s * (VMAX / 32);
  kp.prefill_len = h.prefill_len; kp.rope_delta = h.rope_delta; kp.n_left_pad = h.n_left_pad;
  kp.max_new = h.max_new; kp.min_new = h.min_new; kp.trailing_len = h.trailing_len;
  kp.trailing = h.trailing; kp.tts_pad = h.tts_pad; kp.uniforms = h.uniforms;
  kp.sp_t = h.sp_t; kp.sp_p = h.sp_p;
  kp.nslots = 0;
  return kp;
}

extern "C" int fq3_import_kv(fq3_engine* e, int32_t slot, int32_t layer, const void* k_dev, const void* v_dev, int32_t P,
                             void* stream_) {
  if (!e || !k_dev || !v_dev) return fail(FQ3_ERR_INVALID, "null argument");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  if (P > e->cfg.max_seq_len)
    return fail(FQ3_ERR_TOO_LONG, "Input is too long: prefill has %d tokens but max_seq_len=%d. Use shorter text or shorter reference audio.", P, e->cfg.max_seq_len);
  if (layer < 0 || layer >= e->cfg.talker.num_hidden_layers) return fail(FQ3_ERR_INVALID, "layer out of range");
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nKV = e->cfg.talker.num_key_value_heads, S = e->cfg.max_seq_len;
  const size_t row = (size_t)P * 128 * e->esz, pitch = (size_t)S * 128 * e->esz;
  uint8_t* kd = (uint8_t*)slot_tk(e, slot) + (size_t)layer * nKV * pitch;
  uint8_t* vd = (uint8_t*)slot_tv(e, slot) + (size_t)layer * nKV * pitch;
  if (P > 0) {
    CK(cudaMemcpy2DAsync(kd, pitch, k_dev, row, row, nKV, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpy2DAsync(vd, pitch, v_dev, row, row, nKV, cudaMemcpyDeviceToDevice, stream));
  }
  return 0;
}

extern "C" int fq3_export_kv(fq3_engine* e, int32_t slot, int32_t layer, void* k_dev, void* v_dev, int32_t P, void* stream_) {
  if (!e || !k_dev || !v_dev) return fail(FQ3_ERR_INVALID, "null argument");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  if (P < 0 || P > e->cfg.max_seq_len) return fail(FQ3_ERR_INVALID, "P outside the cache");
  if (layer < 0 || layer >= e->cfg.talker.num_hidden_layers) return fail(FQ3_ERR_INVALID, "layer out of range");
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nKV = e->cfg.talker.num_key_value_heads, S = e->cfg.max_seq_len;
  const size_t row = (size_t)P * 128 * e->esz, pitch = (size_t)S * 128 * e->esz;
  const uint8_t* ks = (const uint8_t*)slot_tk(e, slot) + (size_t)layer * nKV * pitch;
  const uint8_t* vs = (const uint8_t*)slot_tv(e, slot) + (size_t)layer * nKV * pitch;
  if (P > 0) {
    CK(cudaMemcpy2DAsync(k_dev, row, ks, pitch, row, nKV, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpy2DAsync(v_dev, row, vs, pitch, row, nKV, cudaMemcpyDeviceToDevice, stream));
  }
  return 0;
}

extern "C" int fq3_set_generation_state(fq3_engine* e, int32_t slot, int32_t n_left_pad, int32_t rope_delta) {
  if (!e) return fail(FQ3_ERR_INVALID, "null argument");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  e->slots[slot].n_left_pad = n_left_pad;
  e->slots[slot].rope_delta = rope_delta;
  return 0;
}

extern "C" int fq3_talker_step(fq3_engine* e, int32_t slot, const void* embeds_dev, int32_t position, void* hidden_out_dev,
                               void* stream_) {
  if (!e || !embeds_dev || !hidden_out_dev) return fail(FQ3_ERR_INVALID, "null argument");
  if (!e->loaded) return fail(FQ3_ERR_STATE, "weights not loaded");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  if (position < 0 || position >= e->cfg.max_seq_len) return fail(FQ3_ERR_INVALID, "position %d outside the cache", position);
  DevGuard dev_guard(e->dev);
  KParams kp = kp_for_slot(e, slot);
  kp.mode = MODE_TALKER_STEP;
  kp.in_embeds = embeds_dev; kp.hidden_out = hidden_out_dev; kp.position = position;
  kp.dbg_on = e->dbg_on;
  return launch_decode(e, kp, (cudaStream_t)stream_);
}

extern "C" int fq3_predictor_run(fq3_engine* e, int32_t slot, const void* pred_input_dev, const fq3_sampling* sp,
                                 const float* uniforms_dev, int64_t* codes_out_dev, void* stream_) {
  if (!e || !pred_input_dev || !sp || !codes_out_dev) return fail(FQ3_ERR_INVALID, "null argument");
  if (!e->loaded) return fail(FQ3_ERR_STATE, "weights not loaded");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  if (sp->do_sample && !uniforms_dev) return fail(FQ3_ERR_INVALID, "do_sample needs uniforms");
  DevGuard dev_guard(e->dev);
  KParams kp = kp_for_slot(e, slot);
  kp.mode = MODE_PRED_RUN;
  kp.pred_input = pred_input_dev; kp.pred_uniforms = uniforms_dev; kp.codes_out = (long long*)codes_out_dev;
  kp.sp_p = to_sampling(sp);
  kp.dbg_on = e->dbg_on;
  return launch_decode(e, kp, (cudaStream_t)stream_);
}

extern "C" int fq3_sample_logits(fq3_engine* e, const void* logits_dev, int32_t V, const fq3_sampling* sp, float u,
                                 const int64_t* history_dev, int32_t n_hist, int32_t suppress_special, int32_t eos_id,
                                 int32_t suppress_eos, int64_t* token_out_dev, void* stream_) {
  if (!e || !logits_dev || !sp || !token_out_dev) return fail(FQ3_ERR_INVALID, "null argument");
  if (V <= 0 || V > VMAX) return fail(FQ3_ERR_INVALID, "V out of range");
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const Sampling s = to_sampling(sp);
  if (e->bf16)
    sample_kernel<true><<<1, NCT, smem_bytes(), stream>>>(e->kp, logits_dev, V, s, u, (const long long*)history_dev, history_dev ? n_hist : 0, suppress_special, eos_id, suppress_eos, (long long*)token_out_dev);
  else
    sample_kernel<false><<<1, NCT, smem_bytes(), stream>>>(e->kp, logits_dev, V, s, u, (const long long*)history_dev, history_dev ? n_hist : 0, suppress_special, eos_id, suppress_eos, (long long*)token_out_dev);
  e->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int fq3_begin_request(fq3_engine* e, int32_t slot, const fq3_request* rq, const void* past_hidden_dev,
                                 const void* trailing_text_dev, const void* tts_pad_dev, const float* uniforms_dev,
                                 const fq3_sampling* sp_talker, const fq3_sampling* sp_predictor, void* stream_) {
  if (!e || !rq || !past_hidden_dev || !tts_pad_dev || !sp_talker || !sp_predictor) return fail(FQ3_ERR_INVALID, "null argument");
  if (!e->loaded) return fail(FQ3_ERR_STATE, "weights not loaded");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  if (rq->prefill_len > e->cfg.max_seq_len)
    return fail(FQ3_ERR_TOO_LONG, "Input is too long: prefill has %d tokens but max_seq_len=%d. Use shorter text or shorter reference audio.", rq->prefill_len, e->cfg.max_seq_len);
  if ((sp_talker->do_sample || sp_predictor->do_sample) && !uniforms_dev) return fail(FQ3_ERR_INVALID, "sampling needs uniforms");
  if (rq->trailing_len > 0 && !trailing_text_dev) return fail(FQ3_ERR_INVALID, "trailing_len > 0 but no trailing text");
  if (rq->first_token < 0 || rq->first_token >= e->cfg.talker.vocab_size) return fail(FQ3_ERR_INVALID, "first_token out of range");
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  SlotHost& h = e->slots[slot];
  h.prefill_len = rq->prefill_len; h.rope_delta = rq->rope_delta; h.n_left_pad = rq->n_left_pad;
  h.max_new = rq->max_new_tokens; h.min_new = rq->min_new_tokens; h.trailing_len = rq->trailing_len;
  h.trailing = trailing_text_dev; h.tts_pad = tts_pad_dev; h.uniforms = uniforms_dev;
  h.sp_t = to_sampling(sp_talker); h.sp_p = to_sampling(sp_predictor);
  int* st = e->state + 8 * slot;
  float* ph = e->past_hidden + (size_t)slot * HMAX;
  uint32_t* seen = e->seen + (size_t)slot * (VMAX / 32);
  if (e->bf16) set_state_kernel<true><<<4, 256, 0, stream>>>(st, ph, seen, past_hidden_dev, e->cfg.talker.hidden_size, rq->first_token, rq->gen_step);
  else set_state_kernel<false><<<4, 256, 0, stream>>>(st, ph, seen, past_hidden_dev, e->cfg.talker.hidden_size, rq->first_token, rq->gen_step);
  e->launches++;
  CK(cudaGetLastError());
  h.active = true;
  return 0;
}

// batched launch: slots[0..n) become the columns of one pass over the weight tape
static int launch_decode_batch(fq3_engine* e, const int32_t* slots, int n, int n_frames, long long* codes_out_dev,
                               cudaStream_t stream) {
  if (!e->sl_dev) return fail(FQ3_ERR_STATE, "engine was created with max_batch = 1");
  if (n > e->ncta) return fail(FQ3_ERR_INVALID, "%d slots need at least as many CTAs (engine has %d)", n, e->ncta);
  if (e->cfg.has_mtp_projection && !e->kp.mtp_tab)
    return fail(FQ3_ERR_STATE, "batched decode needs the tabulated predictor input projection (unset FQ3_NO_MTP_TABLE)");
  for (int j = 0; j < n; ++j) {
    const int s = slots[j];
    const SlotHost& h = e->slots[s];
    SlotParams& p = e->sl_host[j];
    p.kc = slot_tk(e, s); p.vc = slot_tv(e, s); p.pkc = slot_pk(e, s); p.pvc = slot_pv(e, s);
    p.state = e->state + 8 * s;
    p.past_hidden = e->past_hidden + (size_t)s * HMAX;
    p.seen = e->seen + (size_t)s * (VMAX / 32);
    p.trailing = h.trailing; p.tts_pad = h.tts_pad; p.uniforms = h.uniforms;
    p.codes_out = codes_out_dev + (size_t)j * n_frames * 16;
    p.prefill_len = h.prefill_len; p.rope_delta = h.rope_delta; p.n_left_pad = h.n_left_pad;
    p.max_new = h.max_new; p.min_new = h.min_new; p.trailing_len = h.trailing_len;
    p.sp_t = h.sp_t; p.sp_p = h.sp_p;
  }
  CK(cudaMemcpyAsync(e->sl_dev, e->sl_host, (size_t)n * sizeof(SlotParams), cudaMemcpyHostToDevice, stream));
  CK(cudaMemsetAsync(e->bar, 0, 32768, stream));
  KParams kp = e->kp;
  kp.mode = MODE_FUSED;
  kp.nslots = n;
  kp.sl = e->sl_dev;
  kp.n_frames = n_frames;
  kp.dbg_on = e->dbg_on & 2;
  void* args[] = {(void*)&kp};
  if (e->bf16)
    CK(cudaLaunchCooperativeKernel((const void*)fq3_decode_batch_kernel<true>, dim3(e->ncta), dim3(NTHREADS), args, smem_bytes(), stream));
  else
    CK(cudaLaunchCooperativeKernel((const void*)fq3_decode_batch_kernel<false>, dim3(e->ncta), dim3(NTHREADS), args, smem_bytes(), stream));
  e->launches++;
  return 0;
}

// numerics probe: ONE batched GEMV (the kernel the batched decode path is built from) over a weight segment
extern "C" int fq3_debug_gemv(fq3_engine* e, int32_t stack, int32_t layer, int32_t which, int32_t ncols, const void* x_dev,
                              void* out_dev, void* stream_) {
  if (!e || !x_dev || !out_dev) return fail(FQ3_ERR_INVALID, "null argument");
  if (!e->loaded) return fail(FQ3_ERR_STATE, "weights not loaded");
  if (!e->sl_dev) return fail(FQ3_ERR_STATE, "engine was created with max_batch = 1");
  if (ncols < 1 || ncols > MAXCOL) return fail(FQ3_ERR_INVALID, "ncols out of range");
  const StackDev& S = stack == 0 ? e->kp.t : e->kp.p;
  int sg;
  if (which >= 0 && which < 4) {
    if (layer < 0 || layer >= S.L) return fail(FQ3_ERR_INVALID, "layer out of range");
    sg = S.seg_base + 4 * layer + which;
  } else if (which == 4) {
    sg = S.seg_head + (stack == 0 ? 0 : layer);
  } else {
    return fail(FQ3_ERR_INVALID, "which must be 0..4 (qkv, o, gate/up, down, head)");
  }
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  CK(cudaMemsetAsync(e->bar, 0, 32768, stream));
  KParams kp = e->kp;
  kp.mode = MODE_GEMV_TEST;
  kp.nslots = 1;
  kp.sl = e->sl_dev;
  kp.n_frames = 0;
  kp.dbg_on = 0;
  kp.gt_seg = sg; kp.gt_K = e->segs[sg].K; kp.gt_rows = e->segs[sg].rows; kp.gt_ncols = ncols;
  kp.gt_swiglu = which == 2 ? 1 : 0;
  kp.gt_x = x_dev; kp.gt_out = out_dev;
  void* args[] = {(void*)&kp};
  if (e->bf16)
    CK(cudaLaunchCooperativeKernel((const void*)fq3_decode_batch_kernel<true>, dim3(e->ncta), dim3(NTHREADS), args, smem_bytes(), stream));
  else
    CK(cudaLaunchCooperativeKernel((const void*)fq3_decode_batch_kernel<false>, dim3(e->ncta), dim3(NTHREADS), args, smem_bytes(), stream));
  e->launches++;
  return 0;
}

extern "C" int fq3_decode_chunk(fq3_engine* e, const int32_t* slots, int32_t n_slots, int32_t n_frames,
                                int64_t* codes_out_dev, fq3_chunk_result* res, void* stream_) {
  if (!e || !slots || !codes_out_dev || !res) return fail(FQ3_ERR_INVALID, "null argument");
  if (n_slots <= 0 || n_slots > e->max_batch) return fail(FQ3_ERR_INVALID, "n_slots %d outside [1, max_batch=%d]", n_slots, e->max_batch);
  if (n_frames <= 0) return fail(FQ3_ERR_INVALID, "n_frames must be positive");
  int rc;
  for (int j = 0; j < n_slots; ++j) {
    if ((rc = check_slot(e, slots[j]))) return rc;
    if (!e->slots[slots[j]].active) return fail(FQ3_ERR_STATE, "fq3_begin_request has not been called for slot %d", slots[j]);
    for (int i = 0; i < j; ++i)
      if (slots[i] == slots[j]) return fail(FQ3_ERR_INVALID, "slot %d listed twice", slots[j]);
  }
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_slots == 1) {
    KParams kp = kp_for_slot(e, slots[0]);
    kp.mode = MODE_FUSED;
    kp.n_frames = n_frames;
    kp.codes_out = (long long*)codes_out_dev;
    kp.dbg_on = e->dbg_on & 2;   // timing probes only; layer dumps belong to the step-wise entry points
    if ((rc = launch_decode(e, kp, stream))) return rc;
  } else {
    if ((rc = launch_decode_batch(e, slots, n_slots, n_frames, (long long*)codes_out_dev, stream))) return rc;
  }
  for (int j = 0; j < n_slots; ++j)
    CK(cudaMemcpyAsync(e->state_host + 8 * j, e->state + 8 * slots[j], 32, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  for (int j = 0; j < n_slots; ++j) {
    const int* st = e->state_host + 8 * j;
    res[j].next_token = st[0];
    res[j].total_frames = st[1];
    res[j].finished = st[3];
    res[j].frames_emitted = st[4];
  }
  return 0;
}

extern "C" int fq3_get_past_hidden(fq3_engine* e, int32_t slot, void* dst_dev, void* stream_) {
  if (!e || !dst_dev) return fail(FQ3_ERR_INVALID, "null argument");
  int rc;
  if ((rc = check_slot(e, slot))) return rc;
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const float* ph = e->past_hidden + (size_t)slot * HMAX;
  if (e->bf16) get_hidden_kernel<true><<<4, 256, 0, stream>>>(ph, dst_dev, e->cfg.talker.hidden_size);
  else get_hidden_kernel<false><<<4, 256, 0, stream>>>(ph, dst_dev, e->cfg.talker.hidden_size);
  e->launches++;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int fq3_barrier_test(fq3_engine* e, int32_t n, int32_t kind, void* stream_) {
  if (!e || !e->loaded) return fail(FQ3_ERR_STATE, "weights not loaded");
  DevGuard dev_guard(e->dev);
  KParams kp = e->kp;
  kp.mode = MODE_BARRIER_TEST;
  kp.n_frames = n;
  kp.position = kind;
  return launch_decode(e, kp, (cudaStream_t)stream_);
}

extern "C" int fq3_debug_enable(fq3_engine* e, int32_t on) {
  if (!e) return fail(FQ3_ERR_INVALID, "null argument");
  e->dbg_on = on;
  return 0;
}

extern "C" int fq3_debug_read(fq3_engine* e, int64_t offset, int64_t count, float* host_dst) {
  if (!e || !host_dst) return fail(FQ3_ERR_INVALID, "null argument");
  if (offset < 0 || count < 0 || (size_t)(offset + count) > e->dbg_floats) return fail(FQ3_ERR_INVALID, "debug range outside the buffer (%zu floats)", e->dbg_floats);
  DevGuard dev_guard(e->dev);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(host_dst, e->dbg + offset, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int fq3_tape_bytes(fq3_engine* e, int64_t* talker_step_bytes, int64_t* predictor_frame_bytes) {
  if (!e) return fail(FQ3_ERR_INVALID, "null argument");
  if (talker_step_bytes) *talker_step_bytes = e->talker_step_bytes;
  if (predictor_frame_bytes) *predictor_frame_bytes = e->predictor_frame_bytes;
  return 0;
}

extern "C" int fq3_num_ctas(fq3_engine* e) { return e ? e->ncta : 0; }
extern "C" int64_t fq3_launch_count(fq3_engine* e) { return e ? e->launches : 0; }
extern "C" const char* fq3_last_error(void) { return g_err; }
extern "C" int fq3_max_batch(fq3_engine* e) { return e ? e->max_batch : 0; }
extern "C" const char* fq3_version(void) { return "fq3-b200 0.2.0 (sm_100a)"; }

#include "fq3_prefill.cuh"
