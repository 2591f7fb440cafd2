// fq3_decode.cuh -- the persistent decode kernel (sm_100a).
//
// One cooperative launch runs a whole chunk of codec frames on device: per frame the 15-pass code predictor
// (reference: faster_qwen3_tts/predictor_graph.py:115-167), the 16-row embedding sum (generate.py:163-171), the
// 28-layer talker step (talker_graph.py:97-107,198-214), codec_head, repetition penalty and sampling
// (generate.py:182-197, sampling.py:10-66) and the EOS / max-length control flow (generate.py:149-151,175-177).
// No per-token launch, graph replay or host round trip remains inside a chunk.
//
// Structure of a CTA (one per SM, 288 threads):
//   warp 8      PRODUCER: walks this CTA's slice of the weight "tape" (weights pre-packed at load time into
//               the exact order they are consumed) and streams it with cp.async.bulk (TMA bulk copy, SASS
//               UBLKCP) into a 5 x 32 KB shared-memory ring guarded by full/empty mbarriers.  It never takes
//               part in grid barriers, so HBM keeps streaming while the consumers synchronise / do attention.
//   warps 0..7  CONSUMERS: fp32-accumulate GEMV rows straight out of the ring (conflict-free 16-byte LDS),
//               warp-shuffle reductions, fused epilogues (residual add, SiLU*up, bias), RMSNorm prologues,
//               GQA attention over the KV cache, and block-wide sampling.  Everything that is cheap is
//               computed redundantly in every CTA (norms, sampling, embedding sums) so that only five grid
//               barriers per layer remain.
//
// Numerics follow the reference's eager bf16 / fp32 rounding points (template parameter BF): products of
// dtype-rounded operands are accumulated in fp32 and rounded to the model dtype wherever torch would
// materialise a tensor.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace fq3 {

constexpr int NCW = 8;               // consumer warps
constexpr int NCT = NCW * 32;        // consumer threads
constexpr int NTHREADS = NCT + 32;   // + producer warp
constexpr int NS = 5;                // ring stages
constexpr int STAGE_BYTES = 32768;
constexpr int XS_FLOATS = 6144;      // activation vector(s) feeding the current GEMV; also attention/sampling scratch
constexpr int HMAX = 2048;           // max talker hidden (xin / past_hidden buffers)
constexpr int VMAX = 4096;
constexpr int MAXGRP = 512;          // row groups per CTA
constexpr int MAXSEG = 256;          // GEMV segments
constexpr int SEQMAX = 4096;

enum Mode { MODE_FUSED = 0, MODE_TALKER_STEP = 1, MODE_PRED_RUN = 2, MODE_BARRIER_TEST = 3, MODE_GEMV_TEST = 4 };

struct Grp {           // one row group of one segment, as seen by one CTA (<= 32 rows, full K)
  uint32_t off16;      // tape offset / 16
  int32_t row0;        // first row (index inside the segment)
  uint16_t rows;       // even
  uint16_t m;          // 512-byte row chunks per tile
  uint16_t ntiles;     // K-chunks
  uint16_t pad;
};

struct Sampling {
  int do_sample, top_k;
  float temperature, top_p, penalty;
};

struct StackDev {
  int H, I, L, nH, nKV, V, qd, kd, rep;
  float eps;
  int seg_base;        // segment id of layer 0 / QKV; layer l uses seg_base + 4*l + {0:QKV,1:O,2:GU,3:DN}
  int seg_head;        // first head segment
  const void *ln_in, *ln_post, *qnorm, *knorm, *ln_f;
  void *kc, *vc;       // [L][nKV][S][128] model dtype
  int S;
  const float *cos, *sin;  // [npos][128] fp32
  int npos;
};

struct SlotParams;  // fq3_decode_batch.cuh

struct KParams {
  StackDev t, p;
  int mode, ncta, nseg, seg_mtp;
  const uint8_t* tape;
  const Grp* grps;
  const uint32_t* segtab;       // [cta][nseg] : (begin << 8) | n   (begin relative to this CTA's first group)
  const uint32_t* cta_grp_off;  // [ncta + 1]
  float *X, *X1, *QKV, *ATT, *ACT, *LOGITS;
  int ldX, ldQKV, ldATT, ldACT;
  unsigned* bar;
  const void* t_embed;
  const void* p_embeds;
  const void* mtp_b;
  const void* mtp_tab;   // [ncb][Vp][Hp] = mtp(embeds[i][code]) precomputed at load time (nullptr: project on the fly)
  int has_mtp, ncb, eos;
  int* state;          // [0] token [1] step [2] gen_step [3] finished [4] emitted(last launch)
  float* past_hidden;  // [Ht] fp32 holding dtype-rounded values
  uint32_t* seen;      // [VMAX/32] bitmap of cb0 history (sampling.py:22 unique())
  int prefill_len, rope_delta, n_left_pad, max_new, min_new, trailing_len, max_seq_len;
  const void* trailing;
  const void* tts_pad;
  const float* uniforms;
  Sampling sp_t, sp_p;
  int n_frames;
  long long* codes_out;
  const void* in_embeds;
  void* hidden_out;
  int position;
  const void* pred_input;
  const float* pred_uniforms;
  float* dbg;
  int dbg_on;
  long long dbg_stride_layer;  // floats per layer record
  int pred_pin_layers;         // predictor layers whose weights are streamed with L2 evict_last
  int attn_split;              // talker attention: CTAs per q-head (keys split across them, K/V slices TMA-staged); 0 = off
  int attn_split_min;          // ... used only when at least this many keys are cached (below, one CTA per q-head is faster)
  float* PART;                 // [nH][attn_split][PART_STRIDE] partial attention results (acc[128], max, sum)
  unsigned* attn_cnt;          // [nH] arrival counters of the splits (cleared with the barrier words every launch)
  int mma_tape;                // 1: bf16 tensor-core fragment layout, 0: fp32 row-chunk layout
  // ---- batched decode (fq3_decode_batch.cuh): B request slots share one pass over the weight tape
  int nslots;                  // columns of this launch (0: single-sequence kernel)
  const SlotParams* sl;        // [nslots] per-column request state (device)
  float *XB, *X1B, *QKVB, *LOGB;   // fp32 [MAXCOL][ldX] / [MAXCOL][ldX] / [MAXCOL][ldQKV] / [MAXB][VMAX]
  void *XNB, *ATTB, *ACTB, *PINB;  // model dtype GEMV inputs [MAXCOL][ldX] / [ldATT] / [ldACT] / [HMAX]
  int* TOKB;                   // [MAXB] cb0 token of every column after the talker sampling step
  // MODE_GEMV_TEST: one batched GEMV over segment `gt_seg` (numerics test of the GEMV against a torch reference)
  int gt_seg, gt_K, gt_ncols, gt_rows, gt_swiglu;
  const void* gt_x;            // model dtype [gt_ncols][gt_K]
  void* gt_out;                // fp32 [gt_ncols][gt_rows]  (gt_swiglu: model dtype [gt_ncols][gt_rows / 2])
};

// ------------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  while (!mbar_try_wait(b, parity)) {
  }
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // consumers only

template <bool BF>
__device__ __forceinline__ float rnd(float x) {
  if constexpr (BF)
    return __bfloat162float(__float2bfloat16_rn(x));
  else
    return x;
}
template <bool BF>
__device__ __forceinline__ float ldw(const void* p, size_t i) {  // read-only weight / table element
  if constexpr (BF)
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(p) + i));
  else
    return __ldg(reinterpret_cast<const float*>(p) + i);
}
template <bool BF>
__device__ __forceinline__ void stw(void* p, size_t i, float v) {
  if constexpr (BF)
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(p)[i] = v;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------------------
// shared memory layout
// ------------------------------------------------------------------------------------------------------------
struct __align__(128) Smem {
  uint8_t ring[NS][STAGE_BYTES];
  float xs[XS_FLOATS];
  float xin[2][HMAX];
  float hid[HMAX];
  Grp grp[MAXGRP];
  uint32_t seg[MAXSEG];
  uint64_t full[NS];
  uint64_t empty[NS];
  float red[NCW];
  int hist[256];
  uint32_t seen[VMAX / 32];
  int codes[16];
  int ibc[4];               // integer broadcast slots
  float fbc[4];             // float broadcast slots
  // hand-shake words between the consumer warps and the producer warp: accessed ONLY through flag_ld / flag_st
  // (shared-memory atomics: well-defined without a block barrier, and silent under compute-sanitizer racecheck)
  int stop_flag;            // consumers -> producer
  int prod_done;            // producer -> consumers
  int prod_issued;          // tiles issued by the producer
  int bst[5][32];           // batched kernel: replicated per-column loop state (token, step, gen_step, finished, emitted)
  long long prof[12];       // batched kernel, CTA 0 / thread 0: [0] last clock [1] current category [2..] cycles per category
  int runl[36];             // batched kernel: columns running in the current (sub)frame, ascending; [32] = their number
};

// All dynamic shared memory of the kernel is one Smem; going through this accessor (instead of a reference carried
// in Ctx) lets the compiler prove the address space everywhere: STS/LDS with 32-bit addresses, not generic ST/LD.
extern __shared__ __align__(128) uint8_t fq3_smem_raw[];
__device__ __forceinline__ Smem& SMEM() { return *reinterpret_cast<Smem*>(fq3_smem_raw); }
static_assert(sizeof(Smem) <= 232448, "Smem exceeds the 227 KB per-CTA limit");

__device__ __forceinline__ int flag_ld(int* p) { return atomicAdd(p, 0); }
__device__ __forceinline__ void flag_st(int* p, int v) { atomicExch(p, v); }

struct Ctx {
  const KParams& P;
  int tid, warp, lane;
  uint32_t tile_ctr;   // tiles consumed (identical in every consumer thread)
  unsigned bar_target; // thread 0 only
};

// ------------------------------------------------------------------------------------------------------------
// grid barrier (consumer warps of all CTAs).  The producer warp never waits here.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_sync_v0(Ctx& c) {  // fence + relaxed atomic + fence (cooperative-groups style)
  csync();
  if (c.tid == 0) {
    c.bar_target += (unsigned)c.P.ncta;
    __threadfence();
    atomicAdd(c.P.bar, 1u);
    while (ld_acquire_u32(c.P.bar) < c.bar_target) {
    }
    __threadfence();
  }
  csync();
}
// default: bar.sync orders the CTA's writes before thread 0's release-reduction (cumulativity); pollers acquire.
__device__ __forceinline__ void grid_sync(Ctx& c) {
  csync();
  if (c.tid == 0) {
    c.bar_target += (unsigned)c.P.ncta;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.P.bar), "r"(1u) : "memory");
    while (ld_acquire_u32(c.P.bar) < c.bar_target) {
    }
  }
  csync();
}

// split form: everything issued between grid_arrive() and grid_wait() overlaps the barrier latency -- used for loads
// that do not depend on other CTAs' results of the current phase (cached K/V rows, norm weights)
__device__ __forceinline__ void grid_arrive(Ctx& c) {
  csync();
  if (c.tid == 0) {
    c.bar_target += (unsigned)c.P.ncta;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.P.bar), "r"(1u) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(Ctx& c) {
  if (c.tid == 0) {
    while (ld_acquire_u32(c.P.bar) < c.bar_target) {
    }
  }
  csync();
}

// experimental variants measured by tools/microbench.py (MODE_BARRIER_TEST)
__device__ __forceinline__ void grid_sync_v1(Ctx& c) {  // release-reduction + acquire-poll, no separate fences
  csync();
  if (c.tid == 0) {
    c.bar_target += (unsigned)c.P.ncta;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.P.bar), "r"(1u) : "memory");
    while (ld_acquire_u32(c.P.bar) < c.bar_target) {
    }
  }
  csync();
}
__device__ __forceinline__ void grid_sync_v2(Ctx& c) {  // two-level: 16 group counters (128 B apart) + top counter
  csync();
  if (c.tid == 0) {
    const unsigned ng = 16;
    const unsigned grp = blockIdx.x % ng;
    const unsigned gsize = (c.P.ncta - grp + ng - 1) / ng;
    c.bar_target += 1;  // epoch
    unsigned* gc = c.P.bar + 32 * (1 + grp);
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(gc), "r"(1u) : "memory");
    if (old + 1 == gsize * c.bar_target)
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.P.bar), "r"(1u) : "memory");
    const unsigned want = ng < (unsigned)c.P.ncta ? ng : (unsigned)c.P.ncta;
    while (ld_acquire_u32(c.P.bar) < want * c.bar_target) {
    }
  }
  csync();
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_sync_v3(Ctx& c) {  // relaxed polling, a single acquire fence at the end
  csync();
  if (c.tid == 0) {
    c.bar_target += (unsigned)c.P.ncta;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(c.P.bar), "r"(1u) : "memory");
    while (ld_relaxed_u32(c.P.bar) < c.bar_target) {
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  csync();
}
__device__ __forceinline__ void grid_sync_v4(Ctx& c) {  // last arriver writes one flag per CTA (128 B apart)
  csync();
  if (c.tid == 0) {
    c.bar_target += 1;  // epoch
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(c.P.bar), "r"(1u) : "memory");
    unsigned* flags = c.P.bar + 32;
    if (old + 1 == (unsigned)c.P.ncta * c.bar_target) {
      for (int i = 0; i < c.P.ncta; ++i)
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32 * i), "r"(c.bar_target) : "memory");
    }
    while (ld_acquire_u32(flags + 32 * blockIdx.x) < c.bar_target) {
    }
  }
  csync();
}

__device__ __forceinline__ float block_sum(Ctx& c, float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (c.lane == 0) SMEM().red[c.warp] = v;
  csync();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < NCW; ++w) r += SMEM().red[w];
  csync();
  return r;
}
__device__ __forceinline__ float block_max(Ctx& c, float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (c.lane == 0) SMEM().red[c.warp] = v;
  csync();
  float r = SMEM().red[0];
#pragma unroll
  for (int w = 1; w < NCW; ++w) r = fmaxf(r, SMEM().red[w]);
  csync();
  return r;
}

// timing probe (dbg_on & 2): CTA 0 / thread 0 appends clock64() to the tail of the debug buffer
__device__ __forceinline__ void probe(Ctx& c, int& idx) {
  if ((c.P.dbg_on & 2) && blockIdx.x == 0 && c.tid == 0) {
    long long* ts = reinterpret_cast<long long*>(c.P.dbg);
    if (idx < 4096) ts[idx] = clock64();
  }
  idx++;
}

__device__ __forceinline__ void probe_at(Ctx& c, int idx) {  // fixed slot (frame-level phases: slots 1024..)
#ifdef FQ3_NO_FRAME_PROBES
  return;
#endif
  if ((c.P.dbg_on & 2) && blockIdx.x == 0 && c.tid == 0)
    reinterpret_cast<long long*>(c.P.dbg)[idx] = clock64();
}

// ------------------------------------------------------------------------------------------------------------
// GEMV over one segment: rows of this CTA, streamed from the ring.  x: shared memory, NT vectors of stride xstride.
// Epilogue epi(row0, v0[NT], v1[NT]) is called by lane 0 for each row pair (rows row0, row0+1).
// ------------------------------------------------------------------------------------------------------------
template <bool BF, int NT, class Epi>
__device__ __forceinline__ void gemv_seg(Ctx& c, int seg, const float* x, int xstride, Epi epi) {
  constexpr int EPL = BF ? 8 : 4;  // elements per lane per 16-byte load
  const uint32_t st = SMEM().seg[seg];
  const int gbeg = (int)(st >> 8), gn = (int)(st & 255u);
  for (int gi = 0; gi < gn; ++gi) {
    const Grp g = SMEM().grp[gbeg + gi];
    const int npairs = g.rows >> 1;
    const int m = g.m;
    float acc[4][NT];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[a][t] = 0.f;
    for (int tl = 0; tl < g.ntiles; ++tl) {
      const int stage = (int)(c.tile_ctr % NS);
      const uint32_t par = (c.tile_ctr / NS) & 1u;
      mbar_wait(&SMEM().full[stage], par);
      const uint8_t* tile = SMEM().ring[stage];
      for (int j = 0; j < m; ++j) {
        const int kb = tl * m + j;
        float xv[NT][EPL];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if constexpr (BF) {
            const float4 a = *reinterpret_cast<const float4*>(x + t * xstride + kb * 256 + c.lane * 4);
            const float4 b = *reinterpret_cast<const float4*>(x + t * xstride + kb * 256 + 128 + c.lane * 4);
            xv[t][0] = a.x; xv[t][1] = a.y; xv[t][2] = a.z; xv[t][3] = a.w;
            xv[t][4] = b.x; xv[t][5] = b.y; xv[t][6] = b.z; xv[t][7] = b.w;
          } else {
            const float4 a = *reinterpret_cast<const float4*>(x + t * xstride + kb * 128 + c.lane * 4);
            xv[t][0] = a.x; xv[t][1] = a.y; xv[t][2] = a.z; xv[t][3] = a.w;
          }
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int p = c.warp + NCW * sl;
          if (p < npairs) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = 2 * p + h;
              const uint4 w = *reinterpret_cast<const uint4*>(tile + ((size_t)(r * m + j) * 32 + c.lane) * 16);
              float wf[EPL];
              if constexpr (BF) {
                wf[0] = bf_lo(w.x); wf[1] = bf_hi(w.x); wf[2] = bf_lo(w.y); wf[3] = bf_hi(w.y);
                wf[4] = bf_lo(w.z); wf[5] = bf_hi(w.z); wf[6] = bf_lo(w.w); wf[7] = bf_hi(w.w);
              } else {
                wf[0] = __uint_as_float(w.x); wf[1] = __uint_as_float(w.y);
                wf[2] = __uint_as_float(w.z); wf[3] = __uint_as_float(w.w);
              }
#pragma unroll
              for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int e = 0; e < EPL; ++e) acc[sl * 2 + h][t] = fmaf(wf[e], xv[t][e], acc[sl * 2 + h][t]);
            }
          }
        }
      }
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&SMEM().empty[stage]);
      c.tile_ctr++;
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int p = c.warp + NCW * sl;
      if (p < npairs) {
        float v0[NT], v1[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          float a = acc[sl * 2][t], b = acc[sl * 2 + 1][t];
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          v0[t] = a;
          v1[t] = b;
        }
        if (c.lane == 0) epi(g.row0 + 2 * p, v0, v1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Producer: stream the segments of the program in consumption order.
// ------------------------------------------------------------------------------------------------------------
// ---- split-key talker attention: which cached keys CTA split s of S handles, and how many 64-key ring tiles that is
constexpr int KVT_KEYS = 64;            // keys per ring tile: K rows at byte 0, V rows at byte KVT_VOFF (bf16, 256 B per row)
constexpr int KVT_VOFF = STAGE_BYTES / 2;
constexpr int PART_STRIDE = 132;
struct KvSlice { int j0, n, ntile; };
__device__ __forceinline__ KvSlice kv_slice(int nold, int S, int s) {
  int per = (nold + S - 1) / S;
  per = (per + 7) & ~7;
  const int j0 = min(s * per, nold), j1 = min(j0 + per, nold);
  return KvSlice{j0, j1 - j0, (j1 - j0 + KVT_KEYS - 1) / KVT_KEYS};
}

struct Producer {
  const KParams& P;
  Smem& s;
  uint32_t ctr;
  bool stopped;
  uint64_t pol_first, pol_last;  // L2 eviction policies: stream-once weights vs weights re-read 15x per frame
  __device__ __forceinline__ void seg(int sg, bool keep = false) {
    if (stopped) return;
    const uint32_t st = s.seg[sg];
    const int gbeg = (int)(st >> 8), gn = (int)(st & 255u);
    for (int gi = 0; gi < gn; ++gi) {
      const Grp g = s.grp[gbeg + gi];
      // fp32 tape: rows x m x 512-byte row chunks; bf16 tape: n_mt x G k-groups x 2048-byte fragment blocks
      const uint32_t bytes = P.mma_tape ? (uint32_t)(g.rows & 0xff) * g.m * 2048u : (uint32_t)g.rows * g.m * 512u;
      const uint8_t* src = P.tape + (size_t)g.off16 * 16;
      for (int tl = 0; tl < g.ntiles; ++tl) {
        const int stage = (int)(ctr % NS);
        const uint32_t par = ((ctr / NS) & 1u) ^ 1u;
        while (!mbar_try_wait(&s.empty[stage], par)) {
          if (flag_ld(&s.stop_flag)) {
            stopped = true;
            return;
          }
        }
        mbar_expect_tx(&s.full[stage], bytes);
        bulk_g2s_hint(s.ring[stage], src + (size_t)tl * bytes, bytes, &s.full[stage], keep ? pol_last : pol_first);
        ++ctr;
      }
    }
  }
  // K/V rows of this CTA's key slice of layer l -> ring tiles (issued right behind the layer's QKV weights, so they
  // land while the QKV GEMV and its barrier are still in flight).  Must mirror attention_split().
  __device__ __forceinline__ void kv_tiles(const StackDev& S, int layer, int slot0, int kv_start) {
    if (stopped) return;
    const int Sx = P.attn_split, b = (int)blockIdx.x;
    if (b >= S.nH * Sx) return;
    const int h = b / Sx, sp = b - h * Sx, g = h / S.rep;
    const KvSlice sl = kv_slice(slot0 - kv_start, Sx, sp);
    const size_t row0 = (size_t)(layer * S.nKV + g) * S.S + kv_start + sl.j0;
    const uint8_t* kb = reinterpret_cast<const uint8_t*>(S.kc) + row0 * 256;
    const uint8_t* vb = reinterpret_cast<const uint8_t*>(S.vc) + row0 * 256;
    for (int tl = 0; tl < sl.ntile; ++tl) {
      const uint32_t bytes = (uint32_t)min(KVT_KEYS, sl.n - KVT_KEYS * tl) * 256u;
      const int stage = (int)(ctr % NS);
      const uint32_t par = ((ctr / NS) & 1u) ^ 1u;
      while (!mbar_try_wait(&s.empty[stage], par)) {
        if (flag_ld(&s.stop_flag)) {
          stopped = true;
          return;
        }
      }
      mbar_expect_tx(&s.full[stage], 2 * bytes);
      bulk_g2s(s.ring[stage], kb + (size_t)tl * KVT_KEYS * 256, bytes, &s.full[stage]);
      bulk_g2s(s.ring[stage] + KVT_VOFF, vb + (size_t)tl * KVT_KEYS * 256, bytes, &s.full[stage]);
      ++ctr;
    }
  }
  // kv_slot0 >= 0: talker step at cache slot kv_slot0 with split attention
  __device__ __forceinline__ void stack_layers(const StackDev& S, int keep_layers = 0, int kv_slot0 = -1, int kv_start = 0) {
    for (int l = 0; l < S.L; ++l)
      for (int q = 0; q < 4; ++q) {
        seg(S.seg_base + 4 * l + q, l < keep_layers);
#ifndef FQ3_NO_SPLIT
        if (q == 0 && kv_slot0 >= 0 && P.attn_split > 0 && P.mma_tape && kv_slot0 - kv_start >= P.attn_split_min)
          kv_tiles(S, l, kv_slot0, kv_start);
#endif
      }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Attention for one q-head over the KV cache (transformers eager_attention_forward semantics, GQA by repeat_kv).
// nt tokens (1, or 2 for the predictor prefill), cache slots slot0.., rotary positions rpos0...
// Also applies q_norm/k_norm + RoPE to the new q/k and appends K,V to the cache (talker_graph StaticCache.update).
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__device__ void attention_head(Ctx& c, const StackDev& S, int layer, int h, int nt, int slot0, int rpos0,
                               int kv_start) {
  const KParams& P = c.P;
  float* sc = SMEM().xs;                // scores: [nt][scw]
  const int scw = (nt == 1) ? SEQMAX : 32;
  float* qs = SMEM().xs + SEQMAX;       // [2][128]
  float* ks = qs + 256;              // [2][128]
  float* vs = ks + 256;              // [2][128]
  float* opart = vs + 256;           // [8][128]
  const int g = h / S.rep;
  const size_t esz = BF ? 2 : 4;
  const size_t head_stride = (size_t)S.S * 128;
  uint8_t* kbase = reinterpret_cast<uint8_t*>(S.kc) + ((size_t)(layer * S.nKV + g) * head_stride) * esz;
  uint8_t* vbase = reinterpret_cast<uint8_t*>(S.vc) + ((size_t)(layer * S.nKV + g) * head_stride) * esz;

  // --- a. q/k norm + rope, v copy.  warp 3*t + {0:q,1:k,2:v}; lane owns e, e+32, e+64, e+96
  if (c.warp < 3 * nt) {
    const int t = c.warp / 3, what = c.warp % 3;
    const float* src = P.QKV + (size_t)t * P.ldQKV + (what == 0 ? h * 128 : (what == 1 ? S.qd + g * 128 : S.qd + S.kd + g * 128));
    float v[4], nwv[4], cc[4], sv[4];
    {
      int rp = rpos0 + t;
      rp = rp < 0 ? 0 : (rp >= S.npos ? S.npos - 1 : rp);
      const float* cs = S.cos + (size_t)rp * 128;
      const float* sn = S.sin + (size_t)rp * 128;
      const void* nw = what == 0 ? S.qnorm : S.knorm;
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // all global loads of this step issued back to back
        const int e = c.lane + 32 * i;
        v[i] = __ldcg(src + e);
        nwv[i] = what < 2 ? ldw<BF>(nw, (size_t)layer * 128 + e) : 0.f;
        cc[i] = what < 2 ? __ldg(cs + e) : 0.f;
        sv[i] = what < 2 ? __ldg(sn + e) : 0.f;
      }
    }
    if (what < 2) {
      float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float r = 1.0f / sqrtf(ss / 128.0f + S.eps);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = rnd<BF>(nwv[i] * rnd<BF>(v[i] * r));
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float rot = (i < 2) ? -v[i + 2] : v[i - 2];
        o[i] = rnd<BF>(rnd<BF>(v[i] * rnd<BF>(cc[i])) + rnd<BF>(rot * rnd<BF>(sv[i])));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = o[i];
    }
    float* dst = (what == 0 ? qs : (what == 1 ? ks : vs)) + t * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[c.lane + 32 * i] = v[i];
    if (what > 0 && (h % S.rep) == 0) {
      uint8_t* cb = (what == 1 ? kbase : vbase) + (size_t)(slot0 + t) * 128 * esz;
#pragma unroll
      for (int i = 0; i < 4; ++i) stw<BF>(cb, c.lane + 32 * i, v[i]);
    }
  }
  csync();

  const float scale = 0.08838834764831845f;  // 128^-0.5
  for (int t = 0; t < nt; ++t) {
    const int last = slot0 + t;          // newest key slot visible to token t
    const int nk = last + 1 - kv_start;  // number of visible keys
    const int nold = slot0 - kv_start;   // keys that live in the global cache
    float* sct = sc + t * scw;
    // --- b. scores
    {
      constexpr int LPK = BF ? 16 : 32;  // lanes per key (16 bytes per lane)
      constexpr int KPW = 32 / LPK;      // keys per warp-instruction
      constexpr int EPL = BF ? 8 : 4;
      constexpr int U = 16;
      const int sub = c.lane % LPK, kin = c.lane / LPK;
      float q[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) q[e] = qs[t * 128 + sub * EPL + e];
      for (int base = 0; base < nold; base += NCW * KPW * U) {
        uint4 kv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = base + (u * NCW + c.warp) * KPW + kin;
          if (jj < nold)
            kv[u] = __ldcg(reinterpret_cast<const uint4*>(kbase + ((size_t)(kv_start + jj) * 128) * esz) + sub);
          else
            kv[u] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = base + (u * NCW + c.warp) * KPW + kin;
          float d = 0.f;
          if constexpr (BF) {
            d = fmaf(q[0], bf_lo(kv[u].x), d); d = fmaf(q[1], bf_hi(kv[u].x), d);
            d = fmaf(q[2], bf_lo(kv[u].y), d); d = fmaf(q[3], bf_hi(kv[u].y), d);
            d = fmaf(q[4], bf_lo(kv[u].z), d); d = fmaf(q[5], bf_hi(kv[u].z), d);
            d = fmaf(q[6], bf_lo(kv[u].w), d); d = fmaf(q[7], bf_hi(kv[u].w), d);
          } else {
            d = fmaf(q[0], __uint_as_float(kv[u].x), d); d = fmaf(q[1], __uint_as_float(kv[u].y), d);
            d = fmaf(q[2], __uint_as_float(kv[u].z), d); d = fmaf(q[3], __uint_as_float(kv[u].w), d);
          }
#pragma unroll
          for (int o = LPK / 2; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
          if (sub == 0 && jj < nold) sct[jj] = rnd<BF>(rnd<BF>(d) * scale);
        }
      }
      // new keys (held in shared memory): warp 0, one key per iteration
      if (c.warp == 0) {
        for (int j = 0; j <= t; ++j) {
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) d = fmaf(qs[t * 128 + c.lane + 32 * i], ks[j * 128 + c.lane + 32 * i], d);
#pragma unroll
          for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
          if (c.lane == 0) sct[nold + j] = rnd<BF>(rnd<BF>(d) * scale);
        }
      }
    }
    csync();
    // --- c. softmax (fp32, then rounded to dtype like softmax(..., dtype=float32).to(q.dtype))
    float mx = -INFINITY;
    for (int j = c.tid; j < nk; j += NCT) mx = fmaxf(mx, sct[j]);
    mx = block_max(c, mx);
    float sm = 0.f;
    for (int j = c.tid; j < nk; j += NCT) {
      const float e = expf(sct[j] - mx);
      sct[j] = e;
      sm += e;
    }
    sm = block_sum(c, sm);
    for (int j = c.tid; j < nk; j += NCT) sct[j] = rnd<BF>(sct[j] / sm);
    csync();
    // --- d. P.V : 16-byte loads; bf16: 16 lanes per key (lane owns 8 dims), 2 keys per warp instruction
    {
      constexpr int LPK = BF ? 16 : 32;
      constexpr int KPW = 32 / LPK;
      constexpr int EPL = BF ? 8 : 4;
      constexpr int U = 16;
      const int sub = c.lane % LPK, kin = c.lane / LPK;
      float acc[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
      for (int base = 0; base < nold; base += NCW * KPW * U) {
        uint4 vv[U];
        float pv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = base + (u * NCW + c.warp) * KPW + kin;
          if (jj < nold) {
            vv[u] = __ldcg(reinterpret_cast<const uint4*>(vbase + ((size_t)(kv_start + jj) * 128) * esz) + sub);
            pv[u] = sct[jj];
          } else {
            vv[u] = make_uint4(0, 0, 0, 0);
            pv[u] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if constexpr (BF) {
            acc[0] = fmaf(pv[u], bf_lo(vv[u].x), acc[0]); acc[1] = fmaf(pv[u], bf_hi(vv[u].x), acc[1]);
            acc[2] = fmaf(pv[u], bf_lo(vv[u].y), acc[2]); acc[3] = fmaf(pv[u], bf_hi(vv[u].y), acc[3]);
            acc[4] = fmaf(pv[u], bf_lo(vv[u].z), acc[4]); acc[5] = fmaf(pv[u], bf_hi(vv[u].z), acc[5]);
            acc[6] = fmaf(pv[u], bf_lo(vv[u].w), acc[6]); acc[7] = fmaf(pv[u], bf_hi(vv[u].w), acc[7]);
          } else {
            acc[0] = fmaf(pv[u], __uint_as_float(vv[u].x), acc[0]); acc[1] = fmaf(pv[u], __uint_as_float(vv[u].y), acc[1]);
            acc[2] = fmaf(pv[u], __uint_as_float(vv[u].z), acc[2]); acc[3] = fmaf(pv[u], __uint_as_float(vv[u].w), acc[3]);
          }
        }
      }
      if constexpr (BF) {  // fold the two key halves of the warp: lanes sub and sub+16 own the same dims
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
      }
      if (c.warp == 0 && kin == 0) {
        for (int j = 0; j <= t; ++j) {
          const float pj = sct[nold + j];
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pj, vs[j * 128 + sub * EPL + e], acc[e]);
        }
      }
      if (kin == 0) {
        float* op = opart + c.warp * 128 + sub * EPL;
#pragma unroll
        for (int e = 0; e < EPL; ++e) op[e] = acc[e];
      }
    }
    csync();
    if (c.tid < 128) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < NCW; ++w) o += opart[w * 128 + c.tid];
      P.ATT[(size_t)t * P.ldATT + h * 128 + c.tid] = rnd<BF>(o);
    }
    csync();
  }
}

// ------------------------------------------------------------------------------------------------------------
// Talker attention with the keys of every q-head split over P.attn_split CTAs (bf16 engines, one token).
// CTA b = h * S + s handles q-head h and the s-th slice of the cached keys; the K/V rows of that slice were staged
// into ring tiles by the producer warp (TMA bulk copies issued behind the QKV weights, i.e. before the barrier that
// precedes this function), so scores and P.V run out of shared memory.  Each CTA publishes an un-normalised partial
// (sum_j e^{s_j - m} v_j, m, sum_j e^{s_j - m}); the CTA that arrives last at the head's counter merges the S
// partials in split order -- deterministic whichever CTA that is -- and writes the head's slice of ATT.
// Differences to attention_head(): probabilities are not rounded to bf16 before P.V (higher precision, not lower).
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__device__ void attention_split(Ctx& c, const StackDev& S, int layer, int slot0, int rpos0, int kv_start) {
  const KParams& P = c.P;
  const int Sx = P.attn_split, b = (int)blockIdx.x;
  if (b >= S.nH * Sx) return;  // spare CTAs
  const int h = b / Sx, sp = b - h * Sx, g = h / S.rep;
  float* sc = SMEM().xs;           // scores / exponentials of this slice (+ the new key)
  float* qs = SMEM().xs + 512;     // [128]
  float* ks = qs + 128;            // [128]
  float* vs = ks + 128;            // [128]
  float* opart = vs + 128;         // [8][128]
  const size_t esz = BF ? 2 : 4;
  uint8_t* kbase = reinterpret_cast<uint8_t*>(S.kc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
  uint8_t* vbase = reinterpret_cast<uint8_t*>(S.vc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
  // --- a. q/k norm + rope, v copy: warp 0 q, warp 1 k, warp 2 v; lane owns e, e+32, e+64, e+96
  if (c.warp < 3) {
    const int what = c.warp;
    const float* src = P.QKV + (what == 0 ? h * 128 : (what == 1 ? S.qd + g * 128 : S.qd + S.kd + g * 128));
    float v[4], nwv[4], cc[4], sv[4];
    int rp = rpos0;
    rp = rp < 0 ? 0 : (rp >= S.npos ? S.npos - 1 : rp);
    const float* cs = S.cos + (size_t)rp * 128;
    const float* sn = S.sin + (size_t)rp * 128;
    const void* nw = what == 0 ? S.qnorm : S.knorm;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = c.lane + 32 * i;
      v[i] = __ldcg(src + e);
      nwv[i] = what < 2 ? ldw<BF>(nw, (size_t)layer * 128 + e) : 0.f;
      cc[i] = what < 2 ? __ldg(cs + e) : 0.f;
      sv[i] = what < 2 ? __ldg(sn + e) : 0.f;
    }
    if (what < 2) {
      float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float r = 1.0f / sqrtf(ss / 128.0f + S.eps);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = rnd<BF>(nwv[i] * rnd<BF>(v[i] * r));
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float rot = (i < 2) ? -v[i + 2] : v[i - 2];
        o[i] = rnd<BF>(rnd<BF>(v[i] * rnd<BF>(cc[i])) + rnd<BF>(rot * rnd<BF>(sv[i])));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = o[i];
    }
    float* dst = what == 0 ? qs : (what == 1 ? ks : vs);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[c.lane + 32 * i] = v[i];
    if (what > 0 && sp == 0 && (h % S.rep) == 0) {  // one CTA per kv group appends the new row to the cache
      uint8_t* cb = (what == 1 ? kbase : vbase) + (size_t)slot0 * 128 * esz;
#pragma unroll
      for (int i = 0; i < 4; ++i) stw<BF>(cb, c.lane + 32 * i, v[i]);
      asm volatile("fence.proxy.async.global;" ::: "memory");  // later steps read these rows through the async proxy (TMA)
    }
  }
  csync();
  const KvSlice sl = kv_slice(slot0 - kv_start, Sx, sp);
  const bool has_new = sp == Sx - 1;
  const int nloc = sl.n + (has_new ? 1 : 0);
  const float scale = 0.08838834764831845f;  // 128^-0.5
  const int sub = c.lane & 15, kin = c.lane >> 4;  // 16 lanes per key (16 bytes each), 2 keys per warp instruction
  // --- b. scores out of the staged K rows
  {
    float q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] = qs[sub * 8 + e];
    for (int tl = 0; tl < sl.ntile; ++tl) {
      const uint32_t tc = c.tile_ctr + (uint32_t)tl;
      const int stage = (int)(tc % NS);
      mbar_wait(&SMEM().full[stage], (tc / NS) & 1u);
      const uint8_t* kt = SMEM().ring[stage];
      const int n = min(KVT_KEYS, sl.n - KVT_KEYS * tl);
      for (int j0 = 0; j0 < n; j0 += 2 * NCW) {
        const int jj = j0 + c.warp * 2 + kin;
        float d = 0.f;
        if (jj < n) {
          const uint4 kv = *reinterpret_cast<const uint4*>(kt + (size_t)jj * 256 + sub * 16);
          d = fmaf(q[0], bf_lo(kv.x), d); d = fmaf(q[1], bf_hi(kv.x), d);
          d = fmaf(q[2], bf_lo(kv.y), d); d = fmaf(q[3], bf_hi(kv.y), d);
          d = fmaf(q[4], bf_lo(kv.z), d); d = fmaf(q[5], bf_hi(kv.z), d);
          d = fmaf(q[6], bf_lo(kv.w), d); d = fmaf(q[7], bf_hi(kv.w), d);
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (sub == 0 && jj < n) sc[KVT_KEYS * tl + jj] = rnd<BF>(rnd<BF>(d) * scale);
      }
    }
    if (has_new && c.warp == 0) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) d = fmaf(qs[c.lane + 32 * i], ks[c.lane + 32 * i], d);
#pragma unroll
      for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (c.lane == 0) sc[sl.n] = rnd<BF>(rnd<BF>(d) * scale);
    }
  }
  csync();
  // --- c. slice-local softmax statistics (fp32)
  float mx = -INFINITY;
  for (int j = c.tid; j < nloc; j += NCT) mx = fmaxf(mx, sc[j]);
  mx = block_max(c, mx);
  float sm = 0.f;
  for (int j = c.tid; j < nloc; j += NCT) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sm += e;
  }
  sm = block_sum(c, sm);
  // --- d. un-normalised P.V out of the staged V rows
  {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int tl = 0; tl < sl.ntile; ++tl) {
      const uint32_t tc = c.tile_ctr + (uint32_t)tl;
      const uint8_t* vt = SMEM().ring[(int)(tc % NS)] + KVT_VOFF;
      const int n = min(KVT_KEYS, sl.n - KVT_KEYS * tl);
      for (int j0 = 0; j0 < n; j0 += 2 * NCW) {
        const int jj = j0 + c.warp * 2 + kin;
        if (jj < n) {
          const float pv = sc[KVT_KEYS * tl + jj];
          const uint4 vv = *reinterpret_cast<const uint4*>(vt + (size_t)jj * 256 + sub * 16);
          acc[0] = fmaf(pv, bf_lo(vv.x), acc[0]); acc[1] = fmaf(pv, bf_hi(vv.x), acc[1]);
          acc[2] = fmaf(pv, bf_lo(vv.y), acc[2]); acc[3] = fmaf(pv, bf_hi(vv.y), acc[3]);
          acc[4] = fmaf(pv, bf_lo(vv.z), acc[4]); acc[5] = fmaf(pv, bf_hi(vv.z), acc[5]);
          acc[6] = fmaf(pv, bf_lo(vv.w), acc[6]); acc[7] = fmaf(pv, bf_hi(vv.w), acc[7]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);  // fold the two keys of a warp instruction
    if (has_new && c.warp == 0 && kin == 0) {
      const float pj = sc[sl.n];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vs[sub * 8 + e], acc[e]);
    }
    if (kin == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) opart[c.warp * 128 + sub * 8 + e] = acc[e];
    }
    // hand the ring tiles back to the producer
    __syncwarp();
    if (c.lane == 0)
      for (int tl = 0; tl < sl.ntile; ++tl) mbar_arrive(&SMEM().empty[(int)((c.tile_ctr + (uint32_t)tl) % NS)]);
    c.tile_ctr += (uint32_t)sl.ntile;
  }
  csync();
  float* part = P.PART + ((size_t)h * Sx + sp) * PART_STRIDE;
  if (c.tid < 128) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) o += opart[w * 128 + c.tid];
    part[c.tid] = o;
  }
  if (c.tid == 128) { part[128] = mx; part[129] = sm; }
  // --- e. arrive at the head's counter; the last split merges
  csync();
  if (c.tid == 0) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(P.attn_cnt + h), "r"(1u) : "memory");
    SMEM().ibc[0] = ((old + 1u) % (unsigned)Sx) == 0u ? 1 : 0;
  }
  csync();
  if (SMEM().ibc[0] && c.tid < 128) {
    const float* ph = P.PART + (size_t)h * Sx * PART_STRIDE;
    float M = -INFINITY;
    for (int s2 = 0; s2 < Sx; ++s2) M = fmaxf(M, __ldcg(ph + (size_t)s2 * PART_STRIDE + 128));
    float L = 0.f, o = 0.f;
    for (int s2 = 0; s2 < Sx; ++s2) {
      const float m2 = __ldcg(ph + (size_t)s2 * PART_STRIDE + 128);
      const float w = m2 == -INFINITY ? 0.f : expf(m2 - M);
      L = fmaf(w, __ldcg(ph + (size_t)s2 * PART_STRIDE + 129), L);
      o = fmaf(w, __ldcg(ph + (size_t)s2 * PART_STRIDE + c.tid), o);
    }
    P.ATT[h * 128 + c.tid] = rnd<BF>(o / L);
  }
  csync();
}

// ------------------------------------------------------------------------------------------------------------
// Sampling (sampling.py:32-66 + :10-29), computed redundantly and deterministically by every CTA.
// Returns the token to all consumer threads.
// ------------------------------------------------------------------------------------------------------------
struct SampleArgs {
  const float* logits;  // global fp32 (dtype-rounded values)
  int V;
  Sampling sp;
  float u;
  bool use_penalty;     // repetition penalty over the seen bitmap
  int sup0;             // ids in [sup0, V) except eos are suppressed (V = none)   generate.py:46-50
  bool suppress_eos;
  int eos;
};

__device__ __forceinline__ uint32_t fkey(float f) {  // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <bool BF>
__device__ int sample_block(Ctx& c, const SampleArgs& a) {
  float* lg = SMEM().xs;  // [V]
  const int V = a.V;
  const int sup0 = a.sup0;
  for (int v = c.tid; v < V; v += NCT) {
    float l = __ldcg(a.logits + v);
    if (a.use_penalty && a.sp.penalty != 1.0f && ((SMEM().seen[v >> 5] >> (v & 31)) & 1u))
      l = l > 0.f ? rnd<BF>(l / a.sp.penalty) : rnd<BF>(l * a.sp.penalty);
    if ((v >= sup0 && v != a.eos) || (a.suppress_eos && v == a.eos)) l = -INFINITY;
    lg[v] = l;
  }
  csync();
  if (!a.sp.do_sample) {  // argmax, lowest index among maxima
    float bm = -INFINITY;
    int bi = 0x7fffffff;
    for (int v = c.tid; v < V; v += NCT) {
      const float l = lg[v];
      if (l > bm || (l == bm && v < bi)) { bm = l; bi = v; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, bm, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; }
    }
    if (c.lane == 0) { SMEM().red[c.warp] = bm; SMEM().hist[c.warp] = bi; }
    csync();
    float m = SMEM().red[0];
    int bi2 = SMEM().hist[0];
    for (int w = 1; w < NCW; ++w) {
      const float om = SMEM().red[w];
      const int oi = SMEM().hist[w];
      if (om > m || (om == m && oi < bi2)) { m = om; bi2 = oi; }
    }
    csync();
    return bi2;
  }
  // temperature
  for (int v = c.tid; v < V; v += NCT) lg[v] = rnd<BF>(lg[v] / a.sp.temperature);
  csync();
  // top-k with ties kept: threshold = k-th largest value (sampling.py:54-56)
  if (a.sp.top_k > 0 && a.sp.top_k < V) {
    uint32_t prefix = 0, mask = 0;
    int remaining = a.sp.top_k;
    for (int pass = 0; pass < (BF ? 2 : 4); ++pass) {  // bf16-rounded logits have 16 zero low bits
      const int shift = 24 - 8 * pass;
      SMEM().hist[c.tid] = 0;
      csync();
      for (int v = c.tid; v < V; v += NCT) {
        const uint32_t k = fkey(lg[v]);
        if ((k & mask) == prefix) atomicAdd(&SMEM().hist[(k >> shift) & 255u], 1);
      }
      csync();
      if (c.warp == 0) {  // lane L owns bins 255-8L .. 248-8L (descending)
        int cnt[8], tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { cnt[i] = SMEM().hist[255 - 8 * c.lane - i]; tot += cnt[i]; }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, incl, o);
          if (c.lane >= o) incl += n;
        }
        const int excl = incl - tot;
        if (excl < remaining && incl >= remaining) {
          int run = excl;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (run < remaining && run + cnt[i] >= remaining) {
              SMEM().ibc[0] = 255 - 8 * c.lane - i;
              SMEM().ibc[1] = remaining - run;
            }
            run += cnt[i];
          }
        }
      }
      csync();
      prefix |= ((uint32_t)SMEM().ibc[0]) << shift;
      mask |= 255u << shift;
      remaining = SMEM().ibc[1];
      csync();
    }
    const float kth = fkey_inv(prefix);
    for (int v = c.tid; v < V; v += NCT)
      if (lg[v] < kth) lg[v] = -INFINITY;
    csync();
  }
  // top-p (fp32 semantics, order: value desc then index asc; keep position 0 and every position with cum <= top_p)
  if (a.sp.top_p < 1.0f) {
    float* sl = &SMEM().xin[0][0];                                      // sorted values [V] (xin is free here)
    uint16_t* rk = reinterpret_cast<uint16_t*>(SMEM().xs + VMAX);       // ranks [V]
    for (int v = c.tid; v < V; v += NCT) {
      const float l = lg[v];
      int r = 0;
      for (int w = 0; w < V; ++w) {
        const float o = lg[w];
        r += (o > l || (o == l && w < v)) ? 1 : 0;
      }
      rk[v] = (uint16_t)r;
      sl[r] = l;
    }
    csync();
    if (c.tid == 0) {
      const float m0 = sl[0];
      float S = 0.f;
      for (int i = 0; i < V; ++i) S += expf(sl[i] - m0);
      float cum = 0.f;
      int keep = 1;
      for (int i = 0; i < V; ++i) {
        cum += expf(sl[i] - m0) / S;
        if (i > 0 && !(cum > a.sp.top_p)) keep = i + 1;
        if (cum > a.sp.top_p && i > 0) break;
      }
      SMEM().ibc[2] = keep;
    }
    csync();
    const int keep = SMEM().ibc[2];
    for (int v = c.tid; v < V; v += NCT)
      if ((int)rk[v] >= keep) lg[v] = -INFINITY;
    csync();
  }
  // softmax -> probabilities in model dtype (F.softmax on a dtype tensor)
  float mx = -INFINITY;
  for (int v = c.tid; v < V; v += NCT) mx = fmaxf(mx, lg[v]);
  mx = block_max(c, mx);
  float sm = 0.f;
  for (int v = c.tid; v < V; v += NCT) {
    const float e = expf(lg[v] - mx);
    lg[v] = e;
    sm += e;
  }
  sm = block_sum(c, sm);
  for (int v = c.tid; v < V; v += NCT) lg[v] = rnd<BF>(lg[v] / sm);
  csync();
  // inverse-CDF draw, summation order fixed (oracle/qwen3_tts_oracle.py draw_inverse_cdf)
  const int CH = (V + NCT - 1) / NCT;
  float cs = 0.f;
  for (int j = 0; j < CH; ++j) {
    const int idx = c.tid * CH + j;
    if (idx < V) cs += lg[idx];
  }
  float incl = cs;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if (c.lane >= o) incl += n;
  }
  if (c.lane == 31) SMEM().red[c.warp] = incl;
  if (c.tid == 0) {
    SMEM().ibc[3] = -1;
    SMEM().ibc[0] = 0x7fffffff;
  }
  csync();
  float woff = 0.f, total = 0.f;
  for (int w = 0; w < NCW; ++w) {
    if (w == c.warp) woff = total;
    total += SMEM().red[w];
  }
  incl += woff;
  const float target = a.u * total;
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (c.lane == 0) excl = woff;
  // lowest chunk whose inclusive prefix exceeds the target (prefix sums need not be monotone in fp32; the
  // oracle takes the first such chunk too)
  const bool hit = incl > target;
  if (hit) atomicMin(&SMEM().ibc[0], c.tid);
  csync();
  if (hit && SMEM().ibc[0] == c.tid) {
    float run = excl;
    int pick = -1;
    for (int j = 0; j < CH; ++j) {
      const int idx = c.tid * CH + j;
      if (idx >= V) break;
      run += lg[idx];
      if (run > target && lg[idx] > 0.f) { pick = idx; break; }
    }
    SMEM().ibc[3] = pick;
  }
  csync();
  int tok = SMEM().ibc[3];
  if (tok < 0) {  // rounding left nothing selected: last index with p > 0
    int best = -1;
    for (int v = c.tid; v < V; v += NCT)
      if (lg[v] > 0.f) best = v > best ? v : best;
#pragma unroll
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (c.lane == 0) SMEM().hist[c.warp] = best;
    csync();
    tok = SMEM().hist[0];
    for (int w = 1; w < NCW; ++w) tok = max(tok, SMEM().hist[w]);
    if (tok < 0) tok = 0;
  }
  csync();
  return tok;
}

// ------------------------------------------------------------------------------------------------------------
// Tensor-core GEMV (bf16): the tape holds mma.sync m16n8k16 A-fragments in register order, so one LDS.128 per lane
// feeds one mma (16 rows x 16 k).  The activation vector(s) sit in shared memory as bf16 and enter as the B operand
// (column n = token; for 8-row "HALF" tiles columns 2n / 2n+1 carry the two K halves, rows 0-7 / 8-15 of the tile
// hold the matching halves of the weight rows, and c0 + c3 is the full dot product).  Warps split the k-groups of a
// tile; partial accumulators are combined in shared memory in a fixed order.
//   pre(row, tok) -> float   value fetched BEFORE streaming starts (residual), handed back to epi
//   epi(row, tok, v, vup, aux)   GU tiles: row = pair index, v = gate, vup = up
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(float* d, const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

template <int NT, class Pre, class Epi>
__device__ __forceinline__ void gemv_mma(Ctx& c, int seg, int K, Pre pre, Epi epi) {
  const __nv_bfloat16* xh = reinterpret_cast<const __nv_bfloat16*>(SMEM().xs);
  float* red = SMEM().xs + XS_FLOATS / 2;  // [NCW][2][4][32] partial accumulators
  const uint32_t st = SMEM().seg[seg];
  const int gbeg = (int)(st >> 8), gn = (int)(st & 255u);
  const int gq = c.lane >> 2, t = c.lane & 3;
  for (int gi = 0; gi < gn; ++gi) {
    const Grp g = SMEM().grp[gbeg + gi];
    const int n_mt = g.rows & 0xff, kind = g.rows >> 8, G = g.m;
    int tok, koff;
    bool bvalid;
    if (kind == 1) { tok = gq >> 1; koff = (gq & 1) * (K >> 1); bvalid = gq < 2 * NT; }
    else { tok = gq; koff = 0; bvalid = gq < NT; }
    const __nv_bfloat16* xb = xh + (bvalid ? tok * K + koff : 0) + 16 * t;
    float aux[4] = {0.f, 0.f, 0.f, 0.f};
    if (c.warp < n_mt) {
      if (kind == 0) {
        const int rA = g.row0 + c.warp * 16 + gq;
        if (2 * t < NT) { aux[0] = pre(rA, 2 * t); aux[2] = pre(rA + 8, 2 * t); }
        if (2 * t + 1 < NT) { aux[1] = pre(rA, 2 * t + 1); aux[3] = pre(rA + 8, 2 * t + 1); }
      } else if (kind == 1) {
        if (t < NT) aux[0] = pre(g.row0 + gq, t);
      }
    }
    float acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[a][r] = 0.f;
    for (int tl = 0; tl < g.ntiles; ++tl) {
      const int stage = (int)(c.tile_ctr % NS);
      const uint32_t par = (c.tile_ctr / NS) & 1u;
      mbar_wait(&SMEM().full[stage], par);
      const uint8_t* tile = SMEM().ring[stage];
      for (int qq = c.warp; qq < G; qq += NCW) {
        const int kg = tl * G + qq;
        uint4 blo = make_uint4(0, 0, 0, 0), bhi = make_uint4(0, 0, 0, 0);
        if (bvalid) {
          blo = *reinterpret_cast<const uint4*>(xb + 64 * kg);
          bhi = *reinterpret_cast<const uint4*>(xb + 64 * kg + 8);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt < n_mt) {
            const uint4* A = reinterpret_cast<const uint4*>(tile + ((size_t)(mt * G + qq) * 4) * 512) + c.lane;
            const uint4 a0 = A[0], a1 = A[32], a2 = A[64], a3 = A[96];
            mma_bf16(acc[mt], a0, blo.x, blo.y);
            mma_bf16(acc[mt], a1, blo.z, blo.w);
            mma_bf16(acc[mt], a2, bhi.x, bhi.y);
            mma_bf16(acc[mt], a3, bhi.z, bhi.w);
          }
        }
      }
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&SMEM().empty[stage]);
      c.tile_ctr++;
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
      if (mt < n_mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) red[((c.warp * 2 + mt) * 4 + r) * 32 + c.lane] = acc[mt][r];
    csync();
    if (c.warp < n_mt) {
      const int mt = c.warp;
      float cv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float sm = 0.f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) sm += red[((w * 2 + mt) * 4 + r) * 32 + c.lane];
        cv[r] = sm;
      }
      if (kind == 2) {
        const int pair = g.row0 + mt * 8 + gq;
        if (2 * t < NT) epi(pair, 2 * t, cv[0], cv[2], 0.f);
        if (2 * t + 1 < NT) epi(pair, 2 * t + 1, cv[1], cv[3], 0.f);
      } else if (kind == 0) {
        const int rA = g.row0 + mt * 16 + gq;
        if (2 * t < NT) { epi(rA, 2 * t, cv[0], 0.f, aux[0]); epi(rA + 8, 2 * t, cv[2], 0.f, aux[2]); }
        if (2 * t + 1 < NT) { epi(rA, 2 * t + 1, cv[1], 0.f, aux[1]); epi(rA + 8, 2 * t + 1, cv[3], 0.f, aux[3]); }
      } else {
        if (t < NT) epi(g.row0 + gq, t, cv[0] + cv[3], 0.f, aux[0]);
      }
    }
    csync();
  }
}

// dispatch: bf16 -> tensor-core path on the bf16 staging vector, fp32 -> FMA path on the fp32 staging vector
template <bool BF, bool GU, class Pre, class Epi>
__device__ __forceinline__ void gemv_any(Ctx& c, int seg, int nt, int K, Pre pre, Epi epi) {
  if constexpr (BF) {
    if (nt == 1) gemv_mma<1>(c, seg, K, pre, epi);
    else gemv_mma<2>(c, seg, K, pre, epi);
  } else {
    auto epi2 = [&](int row0, const float* v0, const float* v1) {
      for (int t = 0; t < nt; ++t) {
        if constexpr (GU) {
          epi(row0 >> 1, t, v0[t], v1[t], 0.f);
        } else {
          epi(row0, t, v0[t], 0.f, pre(row0, t));
          epi(row0 + 1, t, v1[t], 0.f, pre(row0 + 1, t));
        }
      }
    };
    if (nt == 1) gemv_seg<false, 1>(c, seg, SMEM().xs, K, epi2);
    else gemv_seg<false, 2>(c, seg, SMEM().xs, K, epi2);
  }
}

// staging vector element store: bf16 array (tensor-core path) or fp32 array (fp32 parity mode), both in s.xs
template <bool BF>
__device__ __forceinline__ void xs_put(Ctx& c, int idx, float v) {
  if constexpr (BF) reinterpret_cast<__nv_bfloat16*>(SMEM().xs)[idx] = __float2bfloat16_rn(v);
  else SMEM().xs[idx] = v;
}
template <bool BF>
__device__ __forceinline__ float xs_get(Ctx& c, int idx) {
  if constexpr (BF) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(SMEM().xs)[idx]);
  else return SMEM().xs[idx];
}

// ------------------------------------------------------------------------------------------------------------
// Predictor-size attention (cache <= 32 slots, i.e. <= 17 keys): EVERY CTA computes all heads redundantly from the
// QKV scratch and the tiny KV cache and writes the result straight into the staging vector that feeds o_proj.
// This removes the attention exchange (ATT round trip) and one grid barrier per layer.  One warp per kv group;
// a lane owns dims [4*lane, 4*lane+4) of every 128-vector; CTA 0 appends the new K/V to the cache.
// ------------------------------------------------------------------------------------------------------------
// NT == 2 is the predictor prefill (slot0 == 0: no cached keys at all); NT == 1 the single-token passes.
// Register budget is 168/thread (9 warps per SM), so K rows and V rows are fetched in two round trips.
// ------------------------------------------------------------------------------------------------------------
// Warp reduce-scatter of 34 per-lane partial sums (2 heads x 17 keys of the predictor attention): instead of a 5-round
// butterfly on every value (170 shuffles, every lane ends with every sum), each round halves the value set -- a lane
// keeps one half and hands the other to its partner -- so 36 shuffles leave every sum on exactly one lane.  The owner's
// value is bit-identical to the butterfly's (same pairing tree, fp32 addition commutes).  own_lane / own_slot: where
// the sum of original index idx ends up.
// ------------------------------------------------------------------------------------------------------------
template <int N, int O>
__device__ __forceinline__ void rs_step(float* a, int lane) {
  constexpr int HH = (N + 1) / 2;
  const bool up = (lane & O) != 0;
#pragma unroll
  for (int i = 0; i < HH; ++i) {
    const float lo = a[i];
    const float hi = (i + HH < N) ? a[i + HH] : 0.f;
    const float send = up ? lo : hi, keep = up ? hi : lo;
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
  }
}
__host__ __device__ constexpr int rs34_lane(int idx) {
  int l = 0, r = idx;
  if (r >= 17) { l |= 16; r -= 17; }
  if (r >= 9) { l |= 8; r -= 9; }
  if (r >= 5) { l |= 4; r -= 5; }
  if (r >= 3) { l |= 2; r -= 3; }
  if (r >= 2) { l |= 1; r -= 2; }
  return l;
}
__host__ __device__ constexpr int rs34_slot(int idx) {
  int r = idx;
  if (r >= 17) r -= 17;
  if (r >= 9) r -= 9;
  if (r >= 5) r -= 5;
  if (r >= 3) r -= 3;
  if (r >= 2) r -= 2;
  return r;
}
// scores sc[2][1][17] (per-lane partials) -> mine[hh] = the full dot product of key `lane` (lanes >= 17: untouched)
template <int HHI = 0, int J = 0>
__device__ __forceinline__ void rs34_gather(const float* a, int lane, float* mine) {
  if constexpr (HHI < 2) {
    constexpr int idx = HHI * 17 + J;
    const float v = __shfl_sync(0xffffffffu, a[rs34_slot(idx)], rs34_lane(idx));
    if (lane == J) mine[HHI] = v;
    if constexpr (J + 1 < 17) rs34_gather<HHI, J + 1>(a, lane, mine);
    else rs34_gather<HHI + 1, 0>(a, lane, mine);
  }
}

template <bool BF>
struct SmallKV {  // cached K/V rows of this warp's kv group, fetched in the shadow of the QKV barrier
  using Raw = typename std::conditional<BF, uint2, float4>::type;
  Raw k[16], v[16];
};
template <bool BF>
__device__ __forceinline__ void small_kv_preload(Ctx& c, const StackDev& S, int layer, int slot0, SmallKV<BF>& pre) {
  using Raw = typename SmallKV<BF>::Raw;
  const size_t esz = BF ? 2 : 4;
  const int g = c.warp < S.nKV ? c.warp : 0;
  const uint8_t* kb = reinterpret_cast<const uint8_t*>(S.kc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
  const uint8_t* vb = reinterpret_cast<const uint8_t*>(S.vc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < slot0) {
      pre.k[j] = __ldcg(reinterpret_cast<const Raw*>(kb + (size_t)j * 128 * esz) + c.lane);
      pre.v[j] = __ldcg(reinterpret_cast<const Raw*>(vb + (size_t)j * 128 * esz) + c.lane);
    } else {
      if constexpr (BF) { pre.k[j] = make_uint2(0, 0); pre.v[j] = make_uint2(0, 0); }
      else { pre.k[j] = make_float4(0, 0, 0, 0); pre.v[j] = make_float4(0, 0, 0, 0); }
    }
  }
}

template <bool BF, int NT>
__device__ void attention_small_all(Ctx& c, const StackDev& S, int layer, int slot0_, int rpos0,
                                    const SmallKV<BF>* pre = nullptr) {
  const KParams& P = c.P;
  constexpr int NOLD = NT == 2 ? 1 : 16;  // cached keys that can exist
  constexpr int MAXK = NT == 2 ? 2 : 17;
  const int slot0 = NT == 2 ? 0 : slot0_;
  const float scale = 0.08838834764831845f;
  const size_t esz = BF ? 2 : 4;
  const int L4 = 4 * c.lane;
  using Raw = typename std::conditional<BF, uint2, float4>::type;
  auto unpack = [](const Raw& r, float* o) {
    if constexpr (BF) { o[0] = bf_lo(r.x); o[1] = bf_hi(r.x); o[2] = bf_lo(r.y); o[3] = bf_hi(r.y); }
    else { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
  };
  auto zero_raw = [](Raw& r) {
    if constexpr (BF) r = make_uint2(0, 0);
    else r = make_float4(0, 0, 0, 0);
  };
  for (int g = c.warp; g < S.nKV; g += NCW) {
    const uint8_t* kb = reinterpret_cast<const uint8_t*>(S.kc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
    const uint8_t* vb = reinterpret_cast<const uint8_t*>(S.vc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
    // ---- round trip 1: cached K rows + everything about the new token(s)
    Raw kraw[NOLD];
#pragma unroll
    for (int j = 0; j < NOLD; ++j) {
      if (pre && NT == 1 && g == c.warp) kraw[j] = pre->k[j];
      else if (j < slot0) kraw[j] = __ldcg(reinterpret_cast<const Raw*>(kb + (size_t)j * 128 * esz) + c.lane);
      else zero_raw(kraw[j]);
    }
    float4 qn4, kn4, cs4[NT], sn4[NT], kr4[NT], vr4[NT], qr4[2][NT];
    {
      float qn[4], kn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qn[i] = ldw<BF>(S.qnorm, (size_t)layer * 128 + L4 + i);
        kn[i] = ldw<BF>(S.knorm, (size_t)layer * 128 + L4 + i);
      }
      qn4 = make_float4(qn[0], qn[1], qn[2], qn[3]);
      kn4 = make_float4(kn[0], kn[1], kn[2], kn[3]);
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      int rp = rpos0 + t;
      rp = rp < 0 ? 0 : (rp >= S.npos ? S.npos - 1 : rp);
      cs4[t] = __ldg(reinterpret_cast<const float4*>(S.cos + (size_t)rp * 128) + c.lane);
      sn4[t] = __ldg(reinterpret_cast<const float4*>(S.sin + (size_t)rp * 128) + c.lane);
      const size_t tb = (size_t)t * P.ldQKV;
      kr4[t] = __ldcg(reinterpret_cast<const float4*>(P.QKV + tb + S.qd + g * 128) + c.lane);
      vr4[t] = __ldcg(reinterpret_cast<const float4*>(P.QKV + tb + S.qd + S.kd + g * 128) + c.lane);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
        qr4[hh][t] = __ldcg(reinterpret_cast<const float4*>(P.QKV + tb + (g * S.rep + (hh < S.rep ? hh : 0)) * 128) + c.lane);
    }
    auto norm_rope = [&](float* v, const float4& w4, const float4& c4, const float4& s4) {
      const float w[4] = {w4.x, w4.y, w4.z, w4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
      float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float r = 1.0f / sqrtf(ss / 128.0f + S.eps);
      float o4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = rnd<BF>(w[i] * rnd<BF>(v[i] * r));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float other = __shfl_xor_sync(0xffffffffu, v[i], 16);
        const float rot = c.lane < 16 ? -other : other;
        o4[i] = rnd<BF>(rnd<BF>(v[i] * rnd<BF>(cc[i])) + rnd<BF>(rot * rnd<BF>(sv[i])));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = o4[i];
    };
    float knew[NT][4], vnew[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      knew[t][0] = kr4[t].x; knew[t][1] = kr4[t].y; knew[t][2] = kr4[t].z; knew[t][3] = kr4[t].w;
      vnew[t][0] = vr4[t].x; vnew[t][1] = vr4[t].y; vnew[t][2] = vr4[t].z; vnew[t][3] = vr4[t].w;
      norm_rope(knew[t], kn4, cs4[t], sn4[t]);
      if (blockIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          stw<BF>(const_cast<uint8_t*>(kb) + (size_t)(slot0 + t) * 128 * esz, L4 + i, knew[t][i]);
          stw<BF>(const_cast<uint8_t*>(vb) + (size_t)(slot0 + t) * 128 * esz, L4 + i, vnew[t][i]);
        }
      }
    }
    // ---- scores of every (head, token) of this group; probabilities stay in registers
    float q[2][NT][4];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        q[hh][t][0] = qr4[hh][t].x; q[hh][t][1] = qr4[hh][t].y; q[hh][t][2] = qr4[hh][t].z; q[hh][t][3] = qr4[hh][t].w;
        norm_rope(q[hh][t], qn4, cs4[t], sn4[t]);
      }
    float sc[2][NT][MAXK];
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      float kf[4] = {0.f, 0.f, 0.f, 0.f};
      if (j < NOLD && j < slot0) unpack(kraw[j < NOLD ? j : 0], kf);
#pragma unroll
      for (int tn = 0; tn < NT; ++tn)
        if (j == slot0 + tn) {
#pragma unroll
          for (int i = 0; i < 4; ++i) kf[i] = knew[tn][i];
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          float d = q[hh][t][0] * kf[0];
          d = fmaf(q[hh][t][1], kf[1], d); d = fmaf(q[hh][t][2], kf[2], d); d = fmaf(q[hh][t][3], kf[3], d);
          sc[hh][t][j] = d;
        }
    }
    float dotk[2] = {0.f, 0.f};   // NT == 1: the full q.k of key `lane` for the two heads (reduce-scatter path)
    if constexpr (NT == 1) {
      float a[34];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int j = 0; j < 17; ++j) a[hh * 17 + j] = sc[hh][0][j];
      rs_step<34, 16>(a, c.lane);
      rs_step<17, 8>(a, c.lane);
      rs_step<9, 4>(a, c.lane);
      rs_step<5, 2>(a, c.lane);
      rs_step<3, 1>(a, c.lane);
      rs34_gather(a, c.lane, dotk);
    } else {
#pragma unroll
      for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int j = 0; j < MAXK; ++j) sc[hh][t][j] += __shfl_xor_sync(0xffffffffu, sc[hh][t][j], o);
    }
    // ---- round trip 2: cached V rows (issued before the softmax arithmetic so the latency overlaps it)
    Raw vraw[NOLD];
#pragma unroll
    for (int j = 0; j < NOLD; ++j) {
      if (pre && NT == 1 && g == c.warp) vraw[j] = pre->v[j];
      else if (j < slot0) vraw[j] = __ldcg(reinterpret_cast<const Raw*>(vb + (size_t)j * 128 * esz) + c.lane);
      else zero_raw(vraw[j]);
    }
    // softmax with the keys distributed over lanes: lane j owns key j (its score, its exponential); max and sum are
    // warp reductions and p_j is broadcast with one shuffle per key.
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int nk = slot0 + t + 1;
        float mx = -INFINITY, mine = -INFINITY;
        if constexpr (NT == 1) {
          mine = (c.lane < nk) ? rnd<BF>(rnd<BF>(dotk[hh]) * scale) : -INFINITY;
          mx = mine;
#pragma unroll
          for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        } else {
#pragma unroll
          for (int j = 0; j < MAXK; ++j) {
            const float sj = j < nk ? rnd<BF>(rnd<BF>(sc[hh][t][j]) * scale) : -INFINITY;
            mx = fmaxf(mx, sj);
            if (j == c.lane) mine = sj;
          }
        }
        const float e = (c.lane < nk) ? (BF ? __expf(mine - mx) : expf(mine - mx)) : 0.f;
        float sm = e;
#pragma unroll
        for (int o = 16; o; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        const float pmine = rnd<BF>(BF ? __fdividef(e, sm) : e / sm);
#pragma unroll
        for (int j = 0; j < MAXK; ++j) sc[hh][t][j] = __shfl_sync(0xffffffffu, pmine, j);
      }
    float o4[2][NT][4];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) o4[hh][t][i] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      float vf[4] = {0.f, 0.f, 0.f, 0.f};
      if (j < NOLD && j < slot0) unpack(vraw[j < NOLD ? j : 0], vf);
#pragma unroll
      for (int tn = 0; tn < NT; ++tn)
        if (j == slot0 + tn) {
#pragma unroll
          for (int i = 0; i < 4; ++i) vf[i] = vnew[tn][i];
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int i = 0; i < 4; ++i) o4[hh][t][i] = fmaf(sc[hh][t][j], vf[i], o4[hh][t][i]);
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
      if (hh < S.rep)
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            xs_put<BF>(c, t * S.qd + (g * S.rep + hh) * 128 + L4 + i, rnd<BF>(o4[hh][t][i]));
  }
  csync();
}

// RMSNorm of a vector (global fp32 or shared fp32) into the staging vector at element offset `off`
constexpr int NORM_E = HMAX / NCT;
// norm weights of one vector into registers (issued in the shadow of a barrier: they are touched once per pass and
// usually miss to HBM, while the activations they scale arrive from L2)
template <bool BF>
__device__ __forceinline__ void norm_wload(Ctx& c, const void* w, size_t woff, int H, float* wv) {
#pragma unroll
  for (int i = 0; i < NORM_E; ++i) {
    const int k = c.tid + i * NCT;
    wv[i] = k < H ? ldw<BF>(w, woff + k) : 0.f;
  }
}
template <bool BF>
__device__ __forceinline__ void norm_stage(Ctx& c, const float* src, bool src_smem, const void* w, size_t woff, int H,
                                           float eps, int off, const float* wpre = nullptr) {
  constexpr int MAXE = NORM_E;
  float v[MAXE], wv[MAXE];
#pragma unroll
  for (int i = 0; i < MAXE; ++i) {
    const int k = c.tid + i * NCT;
    v[i] = 0.f;
    wv[i] = 0.f;
    if (k < H) {
      v[i] = src_smem ? src[k] : __ldcg(src + k);
      wv[i] = wpre ? wpre[i] : ldw<BF>(w, woff + k);
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXE; ++i) ss += v[i] * v[i];
  ss = block_sum(c, ss);
  const float r = 1.0f / sqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXE; ++i) {
    const int k = c.tid + i * NCT;
    if (k < H) xs_put<BF>(c, off + k, rnd<BF>(wv[i] * rnd<BF>(v[i] * r)));
  }
  csync();
}

// ------------------------------------------------------------------------------------------------------------
// One pass through a transformer stack for nt tokens held in X (global) or xin (shared, layer 0).
// On return every CTA holds the final-norm hidden of the LAST token in the staging vector [0..H) and X holds the
// residual stream.
// ------------------------------------------------------------------------------------------------------------
// TALKER selects, at compile time, which attention the instantiation contains: the talker's (one q-head per CTA, or
// keys split over several CTAs with TMA-staged K/V) or the predictor's (every CTA computes all heads of the <= 17-key
// cache redundantly).  Two separate real functions: changing one attention cannot perturb the code of the other pass.
template <bool BF, bool TALKER>
__device__ void run_layers(Ctx& c, const StackDev& S, int nt, int slot0, int rpos0, int kv_start,
                                        bool x0_local, bool dbg) {
  constexpr bool is_talker = TALKER;
  const KParams& P = c.P;
  const bool cta0 = blockIdx.x == 0;
  int pi = 0;
  dbg = dbg && (P.dbg_on & 1);
  auto nopre = [](int, int) { return 0.f; };
  float wnext[NORM_E];  // norm weights of the NEXT norm, fetched while the preceding barrier is in flight
  for (int l = 0; l < S.L; ++l) {
    probe(c, pi);  // 0: layer start
    // ---- P1: input norm + QKV rows
    for (int t = 0; t < nt; ++t) {
      const float* wp = l > 0 ? wnext : nullptr;
      if (l == 0 && x0_local) norm_stage<BF>(c, SMEM().xin[t], true, S.ln_in, (size_t)l * S.H, S.H, S.eps, t * S.H, wp);
      else norm_stage<BF>(c, P.X + (size_t)t * P.ldX, false, S.ln_in, (size_t)l * S.H, S.H, S.eps, t * S.H, wp);
    }
    probe(c, pi);  // 1: after input norm
    gemv_any<BF, false>(c, S.seg_base + 4 * l + 0, nt, S.H, nopre,
                        [&](int row, int t, float v, float, float) { P.QKV[(size_t)t * P.ldQKV + row] = rnd<BF>(v); });
    probe(c, pi);  // 2: after QKV gemv
    const bool small_attn = !TALKER;   // geometry checked by fq3_engine_create (cache <= 32 slots, <= 2 q-heads per kv head)
    const bool kv_pre = small_attn && nt == 1 && S.nKV <= NCW;
    SmallKV<BF> skv;
    grid_arrive(c);
    if (kv_pre) small_kv_preload<BF>(c, S, l, slot0, skv);  // cached keys/values do not depend on this layer's QKV
    grid_wait(c);
    probe(c, pi);  // 3: after B1
    if (dbg && cta0) {
      float* d = P.dbg + (size_t)l * P.dbg_stride_layer;
      for (int t = 0; t < nt; ++t)
        for (int k = c.tid; k < S.qd + 2 * S.kd; k += NCT) d[(size_t)t * (S.qd + 2 * S.kd) + k] = __ldcg(P.QKV + (size_t)t * P.ldQKV + k);
    }
    if constexpr (!TALKER) {
      // ---- P2+P3 fused: redundant small attention straight into the staging vector (no exchange, no barrier)
      if (nt == 1) attention_small_all<BF, 1>(c, S, l, slot0, rpos0, kv_pre ? &skv : nullptr);
      else attention_small_all<BF, 2>(c, S, l, slot0, rpos0);
      probe(c, pi);  // 4
      probe(c, pi);  // 5
    } else {
      // ---- P2: attention.  bf16 talker steps: keys split over attn_split CTAs per q-head, K/V slices TMA-staged;
      //          otherwise one q-head per CTA reading the cache directly
#ifdef FQ3_NO_SPLIT
      const bool split = false;
#else
      const bool split = BF && is_talker && nt == 1 && P.attn_split > 0 && slot0 - kv_start >= P.attn_split_min;
#endif
      if (split) attention_split<BF>(c, S, l, slot0, rpos0, kv_start);
      else
        for (int h = blockIdx.x; h < S.nH; h += gridDim.x) attention_head<BF>(c, S, l, h, nt, slot0, rpos0, kv_start);
      if (!split && is_talker && (int)blockIdx.x >= S.nH && slot0 - kv_start > 64) {
        // idle CTAs pull the NEXT layer's keys/values into L2 (evict_last) so the attention CTAs see L2 latency
        const int ln = (l + 1) % S.L;
        const size_t esz = BF ? 2 : 4;
        const int lines_per_row = (int)(128 * esz / 128);
        const int nk = slot0 - kv_start + (ln == 0 ? 1 : 0);
        const long long total = (long long)S.nKV * nk * lines_per_row * 2;
        const int nidle = (int)gridDim.x - S.nH;
        for (long long i = (long long)(blockIdx.x - S.nH) * NCT + c.tid; i < total; i += (long long)nidle * NCT) {
          const int which = (int)(i & 1);
          long long r = i >> 1;
          const int ln_i = (int)(r % lines_per_row);
          r /= lines_per_row;
          const int j = (int)(r % nk);
          const int g = (int)(r / nk);
          const uint8_t* base = reinterpret_cast<const uint8_t*>(which ? S.vc : S.kc) +
                                (((size_t)(ln * S.nKV + g) * S.S + kv_start + j) * 128) * esz + (size_t)ln_i * 128;
          asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base));
        }
      }
      probe(c, pi);  // 4: after attention
      grid_sync(c);
      probe(c, pi);  // 5: after B2
      // ---- P3: o_proj + residual
      for (int t = 0; t < nt; ++t)
        for (int k = c.tid; k < S.qd; k += NCT) xs_put<BF>(c, t * S.qd + k, __ldcg(P.ATT + (size_t)t * P.ldATT + k));
      csync();
    }
    if (dbg && cta0) {
      float* d = P.dbg + (size_t)l * P.dbg_stride_layer + (size_t)2 * (S.qd + 2 * S.kd);
      for (int k = c.tid; k < nt * S.qd; k += NCT) d[k] = xs_get<BF>(c, k);
    }
    {
      const bool loc = (l == 0 && x0_local);
      gemv_any<BF, false>(
          c, S.seg_base + 4 * l + 1, nt, S.qd,
          [&](int row, int t) { return loc ? SMEM().xin[t][row] : __ldcg(P.X + (size_t)t * P.ldX + row); },
          [&](int row, int t, float v, float, float res) { P.X1[(size_t)t * P.ldX + row] = rnd<BF>(res + rnd<BF>(v)); });
    }
    probe(c, pi);  // 6: after O gemv
    float wpost[NORM_E];
    grid_arrive(c);
    norm_wload<BF>(c, S.ln_post, (size_t)l * S.H, S.H, wpost);
    grid_wait(c);
    probe(c, pi);  // 7: after B3
    // ---- P4: post-attention norm + gate/up rows + SiLU*up
    for (int t = 0; t < nt; ++t)
      norm_stage<BF>(c, P.X1 + (size_t)t * P.ldX, false, S.ln_post, (size_t)l * S.H, S.H, S.eps, t * S.H, wpost);
    if (dbg && cta0) {
      float* d = P.dbg + (size_t)l * P.dbg_stride_layer + (size_t)2 * (S.qd + 2 * S.kd) + 2 * S.qd;
      for (int t = 0; t < nt; ++t)
        for (int k = c.tid; k < S.H; k += NCT) d[(size_t)t * S.H + k] = __ldcg(P.X1 + (size_t)t * P.ldX + k);
    }
    gemv_any<BF, true>(c, S.seg_base + 4 * l + 2, nt, S.H, nopre, [&](int pair, int t, float gv, float uv, float) {
      const float gte = rnd<BF>(gv), up = rnd<BF>(uv);
      const float sl = rnd<BF>(gte / (1.0f + expf(-gte)));
      P.ACT[(size_t)t * P.ldACT + pair] = rnd<BF>(sl * up);
    });
    probe(c, pi);  // 8: after GU gemv
    grid_sync(c);
    probe(c, pi);  // 9: after B4
    // ---- P5: down rows + residual
    for (int t = 0; t < nt; ++t)
      for (int k = c.tid; k < S.I; k += NCT) xs_put<BF>(c, t * S.I + k, __ldcg(P.ACT + (size_t)t * P.ldACT + k));
    csync();
    if (dbg && cta0) {
      float* d = P.dbg + (size_t)l * P.dbg_stride_layer + (size_t)2 * (S.qd + 2 * S.kd) + 2 * S.qd + 2 * S.H;
      for (int k = c.tid; k < nt * S.I; k += NCT) d[k] = xs_get<BF>(c, k);
    }
    gemv_any<BF, false>(
        c, S.seg_base + 4 * l + 3, nt, S.I, [&](int row, int t) { return __ldcg(P.X1 + (size_t)t * P.ldX + row); },
        [&](int row, int t, float v, float, float res) { P.X[(size_t)t * P.ldX + row] = rnd<BF>(res + rnd<BF>(v)); });
    probe(c, pi);  // 10: after DN gemv
    grid_arrive(c);
    if (l + 1 < S.L) norm_wload<BF>(c, S.ln_in, (size_t)(l + 1) * S.H, S.H, wnext);
    else norm_wload<BF>(c, S.ln_f, 0, S.H, wnext);
    grid_wait(c);
    probe(c, pi);  // 11: after B5
    if (dbg && cta0) {
      float* d = P.dbg + (size_t)l * P.dbg_stride_layer + (size_t)2 * (S.qd + 2 * S.kd) + 2 * S.qd + 2 * S.H + 2 * S.I;
      for (int t = 0; t < nt; ++t)
        for (int k = c.tid; k < S.H; k += NCT) d[(size_t)t * S.H + k] = __ldcg(P.X + (size_t)t * P.ldX + k);
    }
  }
  // final norm of the last token -> staging vector [0..H)
  norm_stage<BF>(c, P.X + (size_t)(nt - 1) * P.ldX, false, S.ln_f, 0, S.H, S.eps, 0, wnext);
}

// head GEMV (rows of a [V,H] matrix) on the staging vector -> LOGITS, then grid barrier
template <bool BF>
__device__ __forceinline__ void head_logits(Ctx& c, int seg, int H) {
  const KParams& P = c.P;
  gemv_any<BF, false>(c, seg, 1, H, [](int, int) { return 0.f; },
                      [&](int row, int, float v, float, float) { P.LOGITS[row] = rnd<BF>(v); });
  grid_sync(c);
}

// predictor: 15 passes (predictor_graph.py:115-167).  Inputs: s.xin[0] = past_hidden, s.xin[1] = embed(cb0 token).
// Outputs s.codes[1..15].  u15: 15 uniforms.
template <bool BF>
__device__ void predictor_frame(Ctx& c, const float* u15, bool dbg) {
  const KParams& P = c.P;
  const StackDev& S = P.p;
  const int Ht = P.t.H;
  for (int i = 0; i < P.ncb; ++i) {
    const int nt = (i == 0) ? 2 : 1;
    probe_at(c, 1024 + 8 * i + 0);
    const bool tabled = i > 0 && P.has_mtp && P.mtp_tab != nullptr;
    if (i > 0) {
      const int prev = SMEM().codes[i];  // code sampled by pass i-1
      if (tabled) {  // small_to_mtp_projection(codec_embedding[i-1](prev)) was tabulated when the weights were loaded
        for (int k = c.tid; k < S.H; k += NCT)
          SMEM().xin[0][k] = ldw<BF>(P.mtp_tab, ((size_t)(i - 1) * S.V + prev) * S.H + k);
      } else {
        for (int k = c.tid; k < Ht; k += NCT)
          SMEM().xin[0][k] = ldw<BF>(P.p_embeds, ((size_t)(i - 1) * S.V + prev) * Ht + k);
      }
      csync();
    }
    bool x0_local;
    if (P.has_mtp && !tabled) {
      for (int t = 0; t < nt; ++t)
        for (int k = c.tid; k < Ht; k += NCT) xs_put<BF>(c, t * Ht + k, SMEM().xin[t][k]);
      csync();
      gemv_any<BF, false>(c, P.seg_mtp, nt, Ht, [&](int row, int) { return P.mtp_b ? ldw<BF>(P.mtp_b, row) : 0.f; },
                          [&](int row, int t, float v, float, float b) { P.X[(size_t)t * P.ldX + row] = rnd<BF>(v + b); });
      grid_sync(c);
      x0_local = false;
    } else {
      x0_local = true;
    }
    const int slot0 = (i == 0) ? 0 : i + 1;
    probe_at(c, 1024 + 8 * i + 1);
    run_layers<BF, false>(c, S, nt, slot0, slot0, 0, x0_local, dbg && i == 0);
    probe_at(c, 1024 + 8 * i + 2);
    head_logits<BF>(c, S.seg_head + i, S.H);
    probe_at(c, 1024 + 8 * i + 3);
    SampleArgs sa;
    sa.logits = P.LOGITS; sa.V = S.V; sa.sp = P.sp_p; sa.u = u15 ? __ldg(u15 + i) : 0.f;
    sa.use_penalty = false; sa.sup0 = S.V; sa.suppress_eos = false; sa.eos = -1;
    const int tok = sample_block<BF>(c, sa);
    if (c.tid == 0) SMEM().codes[i + 1] = tok;
    csync();
    probe_at(c, 1024 + 8 * i + 4);
  }
}

// ------------------------------------------------------------------------------------------------------------
// The producer warp's whole life, as one real function: its code generation (counters in registers, no spills) must
// not depend on how much register pressure the consumer code around it creates -- a slow producer slows every phase.
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void producer_main(const KParams& P) {
  Smem& s = SMEM();
  const int lane = (int)(threadIdx.x & 31u);
    if (lane == 0) {
      Producer pr{P, s, 0u, false, 0ull, 0ull};
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pr.pol_first));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pr.pol_last));
      if (P.mode == MODE_BARRIER_TEST) {
      } else if (P.mode == MODE_TALKER_STEP) {
        pr.stack_layers(P.t, 0, P.position, P.n_left_pad);
      } else {
        const int iters = P.mode == MODE_FUSED ? P.n_frames : 1;
        for (int f = 0; f < iters && !pr.stopped; ++f) {
          for (int i = 0; i < P.ncb; ++i) {
            if (P.has_mtp && (i == 0 || P.mtp_tab == nullptr)) pr.seg(P.seg_mtp, true);
            pr.stack_layers(P.p, P.pred_pin_layers);
            pr.seg(P.p.seg_head + i);
          }
          if (P.mode == MODE_FUSED) {
            pr.stack_layers(P.t, 0, P.prefill_len + P.state[1] + f, P.n_left_pad);
            pr.seg(P.t.seg_head);
          }
        }
      }
      flag_st(&s.prod_issued, (int)pr.ctr);
      __threadfence_block();
      flag_st(&s.prod_done, 1);
    }
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__global__ void __launch_bounds__(NTHREADS, 1) fq3_decode_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;

  // one-time setup: per-CTA group / segment tables, barriers
  {
    const uint32_t g0 = __ldg(P.cta_grp_off + cta), g1 = __ldg(P.cta_grp_off + cta + 1);
    for (uint32_t i = tid; i < g1 - g0; i += NTHREADS) s.grp[i] = P.grps[g0 + i];
    for (int i = tid; i < P.nseg; i += NTHREADS) s.seg[i] = __ldg(P.segtab + (size_t)cta * P.nseg + i);
    for (int i = tid; i < VMAX / 32; i += NTHREADS) s.seen[i] = P.mode == MODE_FUSED ? P.seen[i] : 0u;
    if (tid == 0) {
      for (int i = 0; i < NS; ++i) {
        mbar_init(&s.full[i], 1);
        mbar_init(&s.empty[i], NCW);
      }
      s.stop_flag = 0;
      s.prod_done = 0;
      s.prod_issued = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();

  if (warp == NCW) {
    producer_main(P);
  } else {
    // ======================================= CONSUMERS ======================================
    Ctx c{P, tid, warp, lane, 0u, 0u};
    const int Ht = P.t.H;
    if (P.mode == MODE_BARRIER_TEST) {
      for (int i = 0; i < P.n_frames; ++i) {
        if (P.position == 0) grid_sync_v0(c);
        else if (P.position == 1) grid_sync_v1(c);
        else if (P.position == 2) grid_sync_v2(c);
        else if (P.position == 3) grid_sync_v3(c);
        else grid_sync_v4(c);
      }
    } else if (P.mode == MODE_TALKER_STEP) {
      for (int k = tid; k < Ht; k += NCT) s.xin[0][k] = ldw<BF>(P.in_embeds, k);
      csync();
      run_layers<BF, true>(c, P.t, 1, P.position, P.position + P.rope_delta, P.n_left_pad, true, P.dbg_on != 0);
      if (cta == 0)
        for (int k = tid; k < Ht; k += NCT) stw<BF>(P.hidden_out, k, xs_get<BF>(c, k));
    } else if (P.mode == MODE_PRED_RUN) {
      for (int k = tid; k < 2 * Ht; k += NCT) s.xin[k / Ht][k % Ht] = ldw<BF>(P.pred_input, k);
      csync();
      predictor_frame<BF>(c, P.sp_p.do_sample ? P.pred_uniforms : nullptr, P.dbg_on != 0);
      if (cta == 0 && tid < P.ncb) P.codes_out[tid] = (long long)s.codes[tid + 1];
    } else {
      // ---------------- fused frame loop: generate.py:149-199 / streaming.py:106-173 ----------------
      int token = P.state[0], step = P.state[1], gen_step = P.state[2];
      int finished = 0, emitted = 0;
      for (int k = tid; k < Ht; k += NCT) s.hid[k] = P.past_hidden[k];
      csync();
      while (true) {
        if (emitted >= P.n_frames) break;
        if (step >= P.max_new) { finished = 1; break; }
        if (token == P.eos) { finished = 2; break; }
        // predictor input: cat(past_hidden, codec_embedding(token))   generate.py:154-155
        for (int k = tid; k < Ht; k += NCT) {
          s.xin[0][k] = s.hid[k];
          s.xin[1][k] = ldw<BF>(P.t_embed, (size_t)token * Ht + k);
        }
        if (tid == 0) {
          s.codes[0] = token;
          s.seen[token >> 5] |= 1u << (token & 31);
        }
        csync();
        const float* urow = P.uniforms + (size_t)(step + 1) * 16;
        const int pslot = 2048 + 8 * (emitted & 63);
        probe_at(c, pslot + 0);
        predictor_frame<BF>(c, P.sp_p.do_sample ? urow + 1 : nullptr, false);
        probe_at(c, pslot + 1);
        if (cta == 0 && tid < 16) P.codes_out[(size_t)emitted * 16 + tid] = (long long)s.codes[tid];
        emitted++;
        // next talker input: sum of 16 embedding rows + trailing text / tts_pad   generate.py:163-171
        {
          const void* extra = gen_step < P.trailing_len ? P.trailing : P.tts_pad;
          const size_t eoff = gen_step < P.trailing_len ? (size_t)gen_step * Ht : 0;
          for (int k = tid; k < Ht; k += NCT) {
            float sm = ldw<BF>(P.t_embed, (size_t)token * Ht + k);
            for (int i = 0; i < P.ncb; ++i) sm += ldw<BF>(P.p_embeds, ((size_t)i * P.p.V + s.codes[i + 1]) * Ht + k);
            s.xin[0][k] = rnd<BF>(rnd<BF>(sm) + ldw<BF>(extra, eoff + k));
          }
          csync();
        }
        const int pos = P.prefill_len + step;
        if (pos >= P.max_seq_len - 1) { finished = 3; step++; break; }   // generate.py:175-177 (frame already emitted)
        probe_at(c, pslot + 2);
        run_layers<BF, true>(c, P.t, 1, pos, pos + P.rope_delta, P.n_left_pad, true, false);
        probe_at(c, pslot + 3);
        for (int k = tid; k < Ht; k += NCT) s.hid[k] = xs_get<BF>(c, k);   // past_hidden = post-norm hidden (generate.py:198)
        csync();
        head_logits<BF>(c, P.t.seg_head, Ht);
        probe_at(c, pslot + 4);
        SampleArgs sa;
        sa.logits = P.LOGITS; sa.V = P.t.V; sa.sp = P.sp_t; sa.u = P.sp_t.do_sample ? __ldg(urow) : 0.f;
        sa.use_penalty = true; sa.sup0 = P.t.V > 1024 ? P.t.V - 1024 : 0;
        sa.suppress_eos = (step + 1) < P.min_new; sa.eos = P.eos;
        token = sample_block<BF>(c, sa);
        probe_at(c, pslot + 5);
        step++;
        gen_step++;
      }
      if (cta == 0) {
        if (tid == 0) {
          P.state[0] = token; P.state[1] = step; P.state[2] = gen_step; P.state[3] = finished; P.state[4] = emitted;
        }
        for (int k = tid; k < Ht; k += NCT) P.past_hidden[k] = s.hid[k];
        for (int i = tid; i < VMAX / 32; i += NCT) P.seen[i] = s.seen[i];
      }
    }
    // ---- drain: stop the producer and wait for every bulk copy it has in flight
    csync();
    if (tid == 0) {
      flag_st(&s.stop_flag, 1);
      __threadfence_block();
      while (!flag_ld(&s.prod_done)) {
      }
      __threadfence_block();
      const uint32_t issued = (uint32_t)flag_ld(&s.prod_issued);
      for (uint32_t t = c.tile_ctr; t < issued; ++t) mbar_wait(&s.full[t % NS], (t / NS) & 1u);
    }
    csync();
  }
  __syncthreads();
}

}  // namespace fq3
