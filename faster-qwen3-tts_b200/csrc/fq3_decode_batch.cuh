// fq3_decode_batch.cuh -- batched persistent decode kernel (sm_100a): up to 32 request slots share ONE pass over the
// weight tape per step (BASELINE config 4: concurrent requests per GPU; the reference's batch handling is the
// left-padded prompt batch of faster_qwen3_tts/model.py:774-787 with per-row pad counts, talker_graph.py:177-187).
//
// Same tape, same producer warp / TMA ring, same grid barrier and the same per-element arithmetic (rounding points,
// accumulation order inside a warp, cross-warp summation order) as the single-sequence kernel in fq3_decode.cuh: row b
// of a batched launch produces bit-for-bit the codes a single-sequence launch produces for that request (tested).
//
// What changes with B > 1:
//   * activations live in global memory (L2-resident) as [column][K] matrices, column = slot (predictor pass 0: column
//     = token * B + slot).  Every GEMV becomes a skinny GEMM: the activation matrix is the mma.m16n8k16 B operand, read
//     straight from L2 in fragment order (the tape's bf16 column permutation makes a lane's 32 bytes contiguous), up to
//     four n-groups of 8 columns per pass over a ring tile; more than 32 columns (fp32 parity mode: more than 8) replay
//     the segment (the CTA's slice is L2-resident by then).
//   * cheap per-row work that the single-sequence kernel computes redundantly in every CTA is distributed: CTA c
//     serves slot c % B (RMSNorm rows, predictor attention, sampling); talker attention items (slot, q-head) are dealt
//     round-robin to all CTAs.  Results that other CTAs need go through global memory + the grid barrier.
//   * every slot carries its own position, left-pad count, rope delta, trailing-text stream, sampling parameters,
//     uniforms, penalty bitmap and KV caches (SlotParams); slots finish independently (EOS / max_new / max_seq_len).
#pragma once
#include "fq3_decode.cuh"

namespace fq3 {

constexpr int MAXB = 32;          // slots per launch
constexpr int MAXCOL = 2 * MAXB;  // activation columns (predictor pass 0 carries 2 tokens per slot)

struct SlotParams {
  void *kc, *vc;        // talker KV cache of this slot   [L][nKV][S][128]
  void *pkc, *pvc;      // predictor KV cache of this slot [Lp][nKVp][32][128]
  int* state;           // [0] token [1] step [2] gen_step [3] finished [4] emitted(last launch)
  float* past_hidden;   // [HMAX] fp32 holding dtype-rounded values
  uint32_t* seen;       // [VMAX/32] cb0 history bitmap
  const void* trailing;
  const void* tts_pad;
  const float* uniforms;
  long long* codes_out; // [n_frames][16]
  int prefill_len, rope_delta, n_left_pad, max_new, min_new, trailing_len;
  Sampling sp_t, sp_p;
};

enum { BS_TOK = 0, BS_STEP = 1, BS_GEN = 2, BS_FIN = 3, BS_EMIT = 4 };

// phase accounting (dbg_on & 2): CTA 0 / thread 0 charges the clock64 cycles since the last mark to the category that
// was current; dumped to the debug buffer at the end of the launch (tools/batch_bench.py --phases)
enum { PC_OTHER = 0, PC_GEMV = 1, PC_BARRIER = 2, PC_NORM = 3, PC_ATTN = 4, PC_SAMPLE = 5, PC_N = 6 };
__device__ __forceinline__ void pmark(Ctx& c, int cat) {
  if ((c.P.dbg_on & 2) && blockIdx.x == 0 && c.tid == 0) {
    Smem& s = SMEM();
    const long long now = clock64();
    s.prof[2 + (int)s.prof[1]] += now - s.prof[0];
    s.prof[0] = now;
    s.prof[1] = cat;
  }
}
__device__ __forceinline__ void grid_sync_p(Ctx& c, int next_cat) {
  pmark(c, PC_BARRIER);
  grid_sync(c);
  pmark(c, next_cat);
}

// ------------------------------------------------------------------------------------------------------------
// GEMV epilogues (what the single-sequence kernel expresses as lambdas)
// ------------------------------------------------------------------------------------------------------------
enum EpiKind { EP_F32 = 0, EP_RESID = 1, EP_SWIGLU = 2, EP_BIAS = 3 };
struct EpiB {
  int kind;
  float* outf;        // fp32 destination [col][ldo]            (F32 / RESID / BIAS)
  void* outd;         // model-dtype destination [col][ldo]     (SWIGLU)
  int ldo;
  const float* res;   // RESID: residual [col][ldres]
  int ldres;
  const void* bias;   // BIAS: [rows] model dtype or nullptr
};
template <bool BF>
__device__ __forceinline__ float epi_pre(const EpiB& e, int row, int col) {
  if (e.kind == EP_RESID) return __ldcg(e.res + (size_t)col * e.ldres + row);
  if (e.kind == EP_BIAS) return e.bias ? ldw<BF>(e.bias, row) : 0.f;
  return 0.f;
}
template <bool BF>
__device__ __forceinline__ void epi_apply(const EpiB& e, int row, int col, float v, float vup, float aux) {
  if (e.kind == EP_F32) {
    e.outf[(size_t)col * e.ldo + row] = rnd<BF>(v);
  } else if (e.kind == EP_RESID) {
    e.outf[(size_t)col * e.ldo + row] = rnd<BF>(aux + rnd<BF>(v));
  } else if (e.kind == EP_SWIGLU) {
    const float gte = rnd<BF>(v), up = rnd<BF>(vup);
    const float sl = rnd<BF>(gte / (1.0f + expf(-gte)));
    stw<BF>(e.outd, (size_t)col * e.ldo + row, rnd<BF>(sl * up));
  } else {
    e.outf[(size_t)col * e.ldo + row] = rnd<BF>(v + aux);
  }
}

// ------------------------------------------------------------------------------------------------------------
// bf16 tensor-core GEMV over up to 8*NG activation columns (one pass over the segment's ring tiles).
// xg: bf16 [col][ldx] in global memory; columns col0 .. col0+ncols.  Per-column arithmetic identical to gemv_mma.
// ------------------------------------------------------------------------------------------------------------
template <int NG>
__device__ __noinline__ void gemv_mma_b(Ctx& c, int seg, int K, const __nv_bfloat16* __restrict__ xg, int ldx, int col0,
                                        int ncols, const EpiB e) {
  constexpr int NACC = 2 * NG;
  float* red = SMEM().xs;  // [NCW][NACC][4][32] partial accumulators (spills over into xin for NG = 4)
  const uint32_t st = SMEM().seg[seg];
  const int gbeg = (int)(st >> 8), gn = (int)(st & 255u);
  const int gq = c.lane >> 2, t = c.lane & 3;
  const int ngr = (ncols + 7) >> 3;   // n-groups in use (FULL / GU tiles)
  const int ngh = (ncols + 3) >> 2;   // token groups of 4 (HALF tiles)
  for (int gi = 0; gi < gn; ++gi) {
    const Grp g = SMEM().grp[gbeg + gi];
    const int n_mt = g.rows & 0xff, kind = g.rows >> 8, G = g.m;
    const int nacc = kind == 1 ? ngh : n_mt * NG;
    float aux[4] = {0.f, 0.f, 0.f, 0.f};
    if (c.warp < nacc) {
      if (kind == 0) {
        const int mt = c.warp / NG, ng = c.warp % NG;
        const int rA = g.row0 + mt * 16 + gq, cA = ng * 8 + 2 * t;
        if (cA < ncols) { aux[0] = epi_pre<true>(e, rA, col0 + cA); aux[2] = epi_pre<true>(e, rA + 8, col0 + cA); }
        if (cA + 1 < ncols) { aux[1] = epi_pre<true>(e, rA, col0 + cA + 1); aux[3] = epi_pre<true>(e, rA + 8, col0 + cA + 1); }
      } else if (kind == 1) {
        const int tok = c.warp * 4 + t;
        if (tok < ncols) aux[0] = epi_pre<true>(e, g.row0 + gq, col0 + tok);
      }
    }
    // per-lane B-operand row pointers
    const __nv_bfloat16* xb[NACC];
    if (kind != 1) {
#pragma unroll
      for (int ng = 0; ng < NG; ++ng) {
        const int col = ng * 8 + gq;
        xb[ng] = xg + (size_t)(col0 + (col < ncols ? col : 0)) * ldx + 16 * t;
      }
#pragma unroll
      for (int ng = NG; ng < NACC; ++ng) xb[ng] = xb[0];
    } else {
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
        const int tok = a * 4 + (gq >> 1);
        xb[a] = xg + (size_t)(col0 + (tok < ncols ? tok : 0)) * ldx + (gq & 1) * (K >> 1) + 16 * t;
      }
    }
    float acc[NACC][4];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[a][r] = 0.f;
    for (int tl = 0; tl < g.ntiles; ++tl) {
      const int stage = (int)(c.tile_ctr % NS);
      const uint32_t par = (c.tile_ctr / NS) & 1u;
      mbar_wait(&SMEM().full[stage], par);
      const uint8_t* tile = SMEM().ring[stage];
      for (int qq = c.warp; qq < G; qq += NCW) {
        const int kg = tl * G + qq;
        if (kind != 1) {
          uint4 blo[NG], bhi[NG];
#pragma unroll
          for (int ng = 0; ng < NG; ++ng) {
            blo[ng] = __ldcg(reinterpret_cast<const uint4*>(xb[ng] + 64 * kg));
            bhi[ng] = __ldcg(reinterpret_cast<const uint4*>(xb[ng] + 64 * kg + 8));
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt < n_mt) {
              const uint4* A = reinterpret_cast<const uint4*>(tile + ((size_t)(mt * G + qq) * 4) * 512) + c.lane;
              const uint4 a0 = A[0], a1 = A[32], a2 = A[64], a3 = A[96];
#pragma unroll
              for (int ng = 0; ng < NG; ++ng) {
                if (ng < ngr) {
                  mma_bf16(acc[mt * NG + ng], a0, blo[ng].x, blo[ng].y);
                  mma_bf16(acc[mt * NG + ng], a1, blo[ng].z, blo[ng].w);
                  mma_bf16(acc[mt * NG + ng], a2, bhi[ng].x, bhi[ng].y);
                  mma_bf16(acc[mt * NG + ng], a3, bhi[ng].z, bhi[ng].w);
                }
              }
            }
          }
        } else {
          const uint4* A = reinterpret_cast<const uint4*>(tile + ((size_t)qq * 4) * 512) + c.lane;
          const uint4 a0 = A[0], a1 = A[32], a2 = A[64], a3 = A[96];
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            if (a < ngh) {
              const uint4 blo = __ldcg(reinterpret_cast<const uint4*>(xb[a] + 64 * kg));
              const uint4 bhi = __ldcg(reinterpret_cast<const uint4*>(xb[a] + 64 * kg + 8));
              mma_bf16(acc[a], a0, blo.x, blo.y);
              mma_bf16(acc[a], a1, blo.z, blo.w);
              mma_bf16(acc[a], a2, bhi.x, bhi.y);
              mma_bf16(acc[a], a3, bhi.z, bhi.w);
            }
          }
        }
      }
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&SMEM().empty[stage]);
      c.tile_ctr++;
    }
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      if (a < nacc)
#pragma unroll
        for (int r = 0; r < 4; ++r) red[((c.warp * NACC + a) * 4 + r) * 32 + c.lane] = acc[a][r];
    csync();
    if (c.warp < nacc) {
      const int a = c.warp;
      float cv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float sm = 0.f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) sm += red[((w * NACC + a) * 4 + r) * 32 + c.lane];
        cv[r] = sm;
      }
      if (kind == 1) {
        const int tok = a * 4 + t;
        if (tok < ncols) epi_apply<true>(e, g.row0 + gq, col0 + tok, cv[0] + cv[3], 0.f, aux[0]);
      } else {
        const int mt = a / NG, ng = a % NG;
        const int cA = ng * 8 + 2 * t;
        if (kind == 2) {
          const int pair = g.row0 + mt * 8 + gq;
          if (cA < ncols) epi_apply<true>(e, pair, col0 + cA, cv[0], cv[2], 0.f);
          if (cA + 1 < ncols) epi_apply<true>(e, pair, col0 + cA + 1, cv[1], cv[3], 0.f);
        } else {
          const int rA = g.row0 + mt * 16 + gq;
          if (cA < ncols) { epi_apply<true>(e, rA, col0 + cA, cv[0], 0.f, aux[0]); epi_apply<true>(e, rA + 8, col0 + cA, cv[2], 0.f, aux[2]); }
          if (cA + 1 < ncols) { epi_apply<true>(e, rA, col0 + cA + 1, cv[1], 0.f, aux[1]); epi_apply<true>(e, rA + 8, col0 + cA + 1, cv[3], 0.f, aux[3]); }
        }
      }
    }
    csync();
  }
}

// ------------------------------------------------------------------------------------------------------------
// fp32 (parity mode) GEMV over up to 8 activation columns: gemv_seg<false, NT> with x read from global memory.
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void gemv_seg_b(Ctx& c, int seg, const float* __restrict__ xg, int ldx, int col0, int ncols,
                                        const EpiB e) {
  constexpr int NT = 8;
  const uint32_t st = SMEM().seg[seg];
  const int gbeg = (int)(st >> 8), gn = (int)(st & 255u);
  const float* xr[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) xr[t] = xg + (size_t)(col0 + (t < ncols ? t : 0)) * ldx + c.lane * 4;
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
        float xv[NT][4];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(xr[t] + kb * 128));
          xv[t][0] = a.x; xv[t][1] = a.y; xv[t][2] = a.z; xv[t][3] = a.w;
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int p = c.warp + NCW * sl;
          if (p < npairs) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = 2 * p + h;
              const uint4 w = *reinterpret_cast<const uint4*>(tile + ((size_t)(r * m + j) * 32 + c.lane) * 16);
              const float wf[4] = {__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w)};
#pragma unroll
              for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[sl * 2 + h][t] = fmaf(wf[q], xv[t][q], acc[sl * 2 + h][t]);
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
        if (c.lane == 0) {
          const int row0 = g.row0 + 2 * p;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (t < ncols) {
              if (e.kind == EP_SWIGLU) {
                epi_apply<false>(e, row0 >> 1, col0 + t, v0[t], v1[t], 0.f);
              } else {
                epi_apply<false>(e, row0, col0 + t, v0[t], 0.f, epi_pre<false>(e, row0, col0 + t));
                epi_apply<false>(e, row0 + 1, col0 + t, v1[t], 0.f, epi_pre<false>(e, row0 + 1, col0 + t));
              }
            }
          }
        }
      }
    }
  }
}

// column blocks per segment pass: the producer replays the segment once per block
__host__ __device__ __forceinline__ int col_blocks(bool bf, int ncols) { return bf ? (ncols + 31) / 32 : (ncols + 7) / 8; }

template <bool BF>
__device__ __forceinline__ void gemv_b(Ctx& c, int seg, int K, const void* xg, int ldx, int ncols, const EpiB& e) {
  if constexpr (BF) {
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(xg);
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int n = min(32, ncols - c0);
      if (n <= 8) gemv_mma_b<1>(c, seg, K, x, ldx, c0, n, e);
      else if (n <= 16) gemv_mma_b<2>(c, seg, K, x, ldx, c0, n, e);
      else gemv_mma_b<4>(c, seg, K, x, ldx, c0, n, e);
    }
  } else {
    const float* x = reinterpret_cast<const float*>(xg);
    for (int c0 = 0; c0 < ncols; c0 += 8) gemv_seg_b(c, seg, x, ldx, c0, min(8, ncols - c0), e);
  }
}

// ------------------------------------------------------------------------------------------------------------
// RMSNorm of one activation row (values from `prov`) into a model-dtype row of the GEMV input matrix.  Arithmetic and
// summation order of norm_stage(); the writes of a row are shared by `nparts` CTAs (every one of them forms the full
// sum of squares).  xcopy / hcopy: optional fp32 copies of the raw row (residual stream) / of the normalised row.
// ------------------------------------------------------------------------------------------------------------
template <bool BF, class Prov>
__device__ __forceinline__ void norm_row_b(Ctx& c, Prov prov, const void* w, size_t woff, int H, float eps, void* xn,
                                           float* xcopy, float* hcopy, int part, int nparts) {
  float v[NORM_E], wv[NORM_E];
#pragma unroll
  for (int i = 0; i < NORM_E; ++i) {
    const int k = c.tid + i * NCT;
    v[i] = 0.f;
    wv[i] = 0.f;
    if (k < H) {
      v[i] = prov(k);
      wv[i] = ldw<BF>(w, woff + k);
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NORM_E; ++i) ss += v[i] * v[i];
  ss = block_sum(c, ss);
  const float r = 1.0f / sqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < NORM_E; ++i) {
    const int k = c.tid + i * NCT;
    if (k < H && (i % nparts) == part) {
      const float y = rnd<BF>(wv[i] * rnd<BF>(v[i] * r));
      stw<BF>(xn, k, y);
      if (xcopy) xcopy[k] = v[i];
      if (hcopy) hcopy[k] = y;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Talker attention for one (slot, q-head) item: attention_head() with one token, explicit pointers, model-dtype output.
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__device__ void attn_item_b(Ctx& c, const StackDev& S, int layer, int h, const float* __restrict__ qkv, void* kc,
                            void* vc, int slot0, int rpos0, int kv_start, void* att_out) {
  float* sc = SMEM().xs;            // scores [SEQMAX]
  float* qs = SMEM().xs + SEQMAX;   // [128]
  float* ks = qs + 256;
  float* vs = ks + 256;
  float* opart = vs + 256;          // [8][128]
  const int g = h / S.rep;
  const size_t esz = BF ? 2 : 4;
  const size_t head_stride = (size_t)S.S * 128;
  uint8_t* kbase = reinterpret_cast<uint8_t*>(kc) + ((size_t)(layer * S.nKV + g) * head_stride) * esz;
  uint8_t* vbase = reinterpret_cast<uint8_t*>(vc) + ((size_t)(layer * S.nKV + g) * head_stride) * esz;
  if (c.warp < 3) {
    const int what = c.warp;
    const float* src = qkv + (what == 0 ? h * 128 : (what == 1 ? S.qd + g * 128 : S.qd + S.kd + g * 128));
    float v[4], nwv[4], cc[4], sv[4];
    {
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
    if (what > 0 && (h % S.rep) == 0) {
      uint8_t* cb = (what == 1 ? kbase : vbase) + (size_t)slot0 * 128 * esz;
#pragma unroll
      for (int i = 0; i < 4; ++i) stw<BF>(cb, c.lane + 32 * i, v[i]);
      // a later single-sequence launch may read these rows through the async proxy (TMA-staged split attention)
      asm volatile("fence.proxy.async.global;" ::: "memory");
    }
  }
  csync();
  const float scale = 0.08838834764831845f;  // 128^-0.5
  const int nk = slot0 + 1 - kv_start;       // visible keys
  const int nold = slot0 - kv_start;         // keys that live in the global cache
  {
    constexpr int LPK = BF ? 16 : 32;
    constexpr int KPW = 32 / LPK;
    constexpr int EPL = BF ? 8 : 4;
    constexpr int U = 16;
    const int sub = c.lane % LPK, kin = c.lane / LPK;
    float q[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) q[e] = qs[sub * EPL + e];
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
        if (sub == 0 && jj < nold) sc[jj] = rnd<BF>(rnd<BF>(d) * scale);
      }
    }
    if (c.warp == 0) {  // the new key (shared memory)
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) d = fmaf(qs[c.lane + 32 * i], ks[c.lane + 32 * i], d);
#pragma unroll
      for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (c.lane == 0) sc[nold] = rnd<BF>(rnd<BF>(d) * scale);
    }
  }
  csync();
  float mx = -INFINITY;
  for (int j = c.tid; j < nk; j += NCT) mx = fmaxf(mx, sc[j]);
  mx = block_max(c, mx);
  float sm = 0.f;
  for (int j = c.tid; j < nk; j += NCT) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sm += e;
  }
  sm = block_sum(c, sm);
  for (int j = c.tid; j < nk; j += NCT) sc[j] = rnd<BF>(sc[j] / sm);
  csync();
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
          pv[u] = sc[jj];
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
    if constexpr (BF) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (c.warp == 0 && kin == 0) {
      const float pj = sc[nold];
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pj, vs[sub * EPL + e], acc[e]);
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
    stw<BF>(att_out, (size_t)h * 128 + c.tid, rnd<BF>(o));
  }
  csync();
}

// ------------------------------------------------------------------------------------------------------------
// Predictor attention of ONE slot (<= 17 keys): attention_small_all() with explicit pointers, run by one CTA per
// slot; one warp per kv group; K/V rows appended to the slot's cache; output rows in the model-dtype ATT matrix.
//   qkv0 / att0: row of token 0; token t lives `tstride_*` elements further.
// ------------------------------------------------------------------------------------------------------------
template <bool BF, int NT>
__device__ void attn_small_b(Ctx& c, const StackDev& S, int layer, int slot0_, int rpos0, const float* __restrict__ qkv0,
                             size_t tstride_qkv, void* pkc, void* pvc, void* att0, size_t tstride_att) {
  constexpr int NOLD = NT == 2 ? 1 : 16;
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
    uint8_t* kb = reinterpret_cast<uint8_t*>(pkc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
    uint8_t* vb = reinterpret_cast<uint8_t*>(pvc) + ((size_t)(layer * S.nKV + g) * S.S * 128) * esz;
    Raw kraw[NOLD];
#pragma unroll
    for (int j = 0; j < NOLD; ++j) {
      if (j < slot0) kraw[j] = __ldcg(reinterpret_cast<const Raw*>(kb + (size_t)j * 128 * esz) + c.lane);
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
      const float* row = qkv0 + (size_t)t * tstride_qkv;
      kr4[t] = __ldcg(reinterpret_cast<const float4*>(row + S.qd + g * 128) + c.lane);
      vr4[t] = __ldcg(reinterpret_cast<const float4*>(row + S.qd + S.kd + g * 128) + c.lane);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
        qr4[hh][t] = __ldcg(reinterpret_cast<const float4*>(row + (g * S.rep + (hh < S.rep ? hh : 0)) * 128) + c.lane);
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
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        stw<BF>(kb + (size_t)(slot0 + t) * 128 * esz, L4 + i, knew[t][i]);
        stw<BF>(vb + (size_t)(slot0 + t) * 128 * esz, L4 + i, vnew[t][i]);
      }
    }
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
    float dotk[2] = {0.f, 0.f};   // NT == 1: reduce-scatter of the 34 partial dot products (fq3_decode.cuh: rs_step)
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
    Raw vraw[NOLD];
#pragma unroll
    for (int j = 0; j < NOLD; ++j) {
      if (j < slot0) vraw[j] = __ldcg(reinterpret_cast<const Raw*>(vb + (size_t)j * 128 * esz) + c.lane);
      else zero_raw(vraw[j]);
    }
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
            stw<BF>(att0, (size_t)t * tstride_att + (g * S.rep + hh) * 128 + L4 + i, rnd<BF>(o4[hh][t][i]));
  }
  csync();
}

// ------------------------------------------------------------------------------------------------------------
// per-CTA view of a batched launch
// ------------------------------------------------------------------------------------------------------------
struct BView {
  int B;       // columns (slots) of this launch
  int b;       // the slot this CTA serves (cta % B)
  int rank;    // rank of this CTA among the CTAs serving slot b
  int gsz;     // CTAs serving slot b
};

// One pass through a transformer stack for nt tokens per slot (columns t*B + b).  On entry the layer-0 input norm has
// been written (XNB) together with the raw rows (XB) -- both published by the barrier this function starts with.  On
// return XNB rows [0,B) hold the final-norm hidden of the last token of every running slot (published), hcopy: the
// talker's past_hidden copy.
template <bool BF, bool TALKER>
__device__ void stack_b(Ctx& c, const StackDev& S, const BView& v, int nt, uint32_t run, int pass_slot0) {
  const KParams& P = c.P;
  const int B = v.B, ncols = nt * B;
  const bool mine = (run >> v.b) & 1u;
  const int nparts = min(v.gsz, S.H / NCT);
  const SlotParams& me = P.sl[v.b];
  for (int l = 0; l < S.L; ++l) {
    if (l > 0) {
      if (mine && v.rank < nparts)
        for (int t = 0; t < nt; ++t) {
          const int col = t * B + v.b;
          const float* src = P.XB + (size_t)col * P.ldX;
          norm_row_b<BF>(c, [&](int k) { return __ldcg(src + k); }, S.ln_in, (size_t)l * S.H, S.H, S.eps,
                         reinterpret_cast<uint8_t*>(P.XNB) + (size_t)col * P.ldX * (BF ? 2 : 4), nullptr, nullptr, v.rank, nparts);
        }
    }
    grid_sync_p(c, PC_GEMV);
    // ---- QKV rows
    {
      EpiB e{EP_F32, P.QKVB, nullptr, P.ldQKV, nullptr, 0, nullptr};
      gemv_b<BF>(c, S.seg_base + 4 * l + 0, S.H, P.XNB, P.ldX, ncols, e);
    }
    grid_sync_p(c, PC_ATTN);
    // ---- attention
    if constexpr (TALKER) {
      // items (running slot, q-head) dealt round-robin to the CTAs
      const int nrun = SMEM().runl[32];
      for (int idx = blockIdx.x; idx < nrun * S.nH; idx += gridDim.x) {
        const int b = SMEM().runl[idx / S.nH], h = idx % S.nH;
        const SlotParams& sp = P.sl[b];
        const int pos = sp.prefill_len + SMEM().bst[BS_STEP][b];
        attn_item_b<BF>(c, S, l, h, P.QKVB + (size_t)b * P.ldQKV, sp.kc, sp.vc, pos, pos + sp.rope_delta, sp.n_left_pad,
                        reinterpret_cast<uint8_t*>(P.ATTB) + (size_t)b * P.ldATT * (BF ? 2 : 4));
      }
    } else {
      if (mine && v.rank == 0) {
        const float* q0 = P.QKVB + (size_t)v.b * P.ldQKV;
        uint8_t* a0 = reinterpret_cast<uint8_t*>(P.ATTB) + (size_t)v.b * P.ldATT * (BF ? 2 : 4);
        if (nt == 1) attn_small_b<BF, 1>(c, S, l, pass_slot0, pass_slot0, q0, 0, me.pkc, me.pvc, a0, 0);
        else attn_small_b<BF, 2>(c, S, l, 0, 0, q0, (size_t)B * P.ldQKV, me.pkc, me.pvc, a0, (size_t)B * P.ldATT);
      }
    }
    grid_sync_p(c, PC_GEMV);
    // ---- o_proj + residual
    {
      EpiB e{EP_RESID, P.X1B, nullptr, P.ldX, P.XB, P.ldX, nullptr};
      gemv_b<BF>(c, S.seg_base + 4 * l + 1, S.qd, P.ATTB, P.ldATT, ncols, e);
    }
    grid_sync_p(c, PC_NORM);
    // ---- post-attention norm
    if (mine && v.rank < nparts)
      for (int t = 0; t < nt; ++t) {
        const int col = t * B + v.b;
        const float* src = P.X1B + (size_t)col * P.ldX;
        norm_row_b<BF>(c, [&](int k) { return __ldcg(src + k); }, S.ln_post, (size_t)l * S.H, S.H, S.eps,
                       reinterpret_cast<uint8_t*>(P.XNB) + (size_t)col * P.ldX * (BF ? 2 : 4), nullptr, nullptr, v.rank, nparts);
      }
    grid_sync_p(c, PC_GEMV);
    // ---- gate/up + SiLU*up
    {
      EpiB e{EP_SWIGLU, nullptr, P.ACTB, P.ldACT, nullptr, 0, nullptr};
      gemv_b<BF>(c, S.seg_base + 4 * l + 2, S.H, P.XNB, P.ldX, ncols, e);
    }
    grid_sync_p(c, PC_GEMV);
    // ---- down + residual
    {
      EpiB e{EP_RESID, P.XB, nullptr, P.ldX, P.X1B, P.ldX, nullptr};
      gemv_b<BF>(c, S.seg_base + 4 * l + 3, S.I, P.ACTB, P.ldACT, ncols, e);
    }
    grid_sync_p(c, PC_NORM);
  }
  // final norm of the last token -> XNB row b (head GEMV input); talker: also the slot's past_hidden
  if (mine && v.rank < nparts) {
    const int col = (nt - 1) * B + v.b;
    const float* src = P.XB + (size_t)col * P.ldX;
    norm_row_b<BF>(c, [&](int k) { return __ldcg(src + k); }, S.ln_f, 0, S.H, S.eps,
                   reinterpret_cast<uint8_t*>(P.XNB) + (size_t)v.b * P.ldX * (BF ? 2 : 4), nullptr,
                   TALKER ? me.past_hidden : nullptr, v.rank, nparts);
  }
  grid_sync_p(c, PC_GEMV);
}

// ------------------------------------------------------------------------------------------------------------
// producer warp of the batched kernel: the same tape walk as producer_main(), every segment replayed once per
// column block (see gemv_b)
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void producer_batch_main(const KParams& P) {
  Smem& s = SMEM();
  const int lane = (int)(threadIdx.x & 31u);
  if (lane == 0) {
    Producer pr{P, s, 0u, false, 0ull, 0ull};
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pr.pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pr.pol_last));
    const bool bf = P.mma_tape != 0;
    const int nb1 = col_blocks(bf, P.nslots), nb2 = col_blocks(bf, 2 * P.nslots);
    auto rep = [&](int sg, int n, bool keep) {
      for (int i = 0; i < n; ++i) pr.seg(sg, keep);
    };
    if (P.mode == MODE_GEMV_TEST) rep(P.gt_seg, col_blocks(bf, P.gt_ncols), false);
    for (int f = 0; P.mode != MODE_GEMV_TEST && f < P.n_frames && !pr.stopped; ++f) {
      if (P.has_mtp) rep(P.seg_mtp, nb2, true);
      for (int i = 0; i < P.ncb; ++i) {
        for (int l = 0; l < P.p.L; ++l)
          for (int q = 0; q < 4; ++q) rep(P.p.seg_base + 4 * l + q, i == 0 ? nb2 : nb1, l < P.pred_pin_layers);
        rep(P.p.seg_head + i, nb1, false);
      }
      for (int l = 0; l < P.t.L; ++l)
        for (int q = 0; q < 4; ++q) rep(P.t.seg_base + 4 * l + q, nb1, false);
      rep(P.t.seg_head, nb1, false);
    }
    flag_st(&s.prod_issued, (int)pr.ctr);
    __threadfence_block();
    flag_st(&s.prod_done, 1);
  }
}

// ------------------------------------------------------------------------------------------------------------
// the batched kernel: generate.py:149-199 / streaming.py:106-173 for every slot of the launch, lock-step per frame
// ------------------------------------------------------------------------------------------------------------
template <bool BF>
__global__ void __launch_bounds__(NTHREADS, 1) fq3_decode_batch_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x;
  const int B = P.nslots;
  BView v;
  v.B = B;
  v.b = cta % B;
  v.rank = cta / B;
  v.gsz = ((int)gridDim.x - v.b + B - 1) / B;
  const SlotParams& me = P.sl[v.b];
  {
    const uint32_t g0 = __ldg(P.cta_grp_off + cta), g1 = __ldg(P.cta_grp_off + cta + 1);
    for (uint32_t i = tid; i < g1 - g0; i += NTHREADS) s.grp[i] = P.grps[g0 + i];
    for (int i = tid; i < P.nseg; i += NTHREADS) s.seg[i] = __ldg(P.segtab + (size_t)cta * P.nseg + i);
    for (int i = tid; i < VMAX / 32; i += NTHREADS) s.seen[i] = P.mode == MODE_GEMV_TEST ? 0u : me.seen[i];
    if (tid < B && P.mode != MODE_GEMV_TEST) {
      const int* st = P.sl[tid].state;
      s.bst[BS_TOK][tid] = st[0];
      s.bst[BS_STEP][tid] = st[1];
      s.bst[BS_GEN][tid] = st[2];
      s.bst[BS_FIN][tid] = 0;
      s.bst[BS_EMIT][tid] = 0;
    }
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
    producer_batch_main(P);
  } else {
    Ctx c{P, tid, warp, lane, 0u, 0u};
    const int Ht = P.t.H;
    const size_t esz = BF ? 2 : 4;
    const StackDev& Sp = P.p;
    if (tid == 0) {
      for (int i = 0; i < 12; ++i) s.prof[i] = 0;
      s.prof[0] = clock64();
    }
    const int npp = min(v.gsz, Sp.H / NCT);   // CTAs sharing the writes of a predictor norm row
    const int npt = min(v.gsz, Ht / NCT);     // ... of a talker norm row
    if (P.mode == MODE_GEMV_TEST) {
      EpiB e{P.gt_swiglu ? EP_SWIGLU : EP_F32, reinterpret_cast<float*>(P.gt_out), P.gt_out,
             P.gt_swiglu ? P.gt_rows / 2 : P.gt_rows, nullptr, 0, nullptr};
      gemv_b<BF>(c, P.gt_seg, P.gt_K, P.gt_x, P.gt_K, P.gt_ncols, e);
      grid_sync(c);
    }
    while (P.mode != MODE_GEMV_TEST) {
      // ---- which slots run this frame (identical decision in every CTA)
      if (warp == 0) {
        bool r = false;
        if (lane < B && s.bst[BS_FIN][lane] == 0 && s.bst[BS_EMIT][lane] < P.n_frames) {
          if (s.bst[BS_STEP][lane] >= P.sl[lane].max_new) s.bst[BS_FIN][lane] = 1;
          else if (s.bst[BS_TOK][lane] == P.eos) s.bst[BS_FIN][lane] = 2;
          else r = true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, r);
        if (lane == 0) s.ibc[0] = (int)m;
      }
      csync();
      const uint32_t run = (uint32_t)s.ibc[0];
      csync();
      if (run == 0u) break;
      const bool mine = (run >> v.b) & 1u;
      const int token = s.bst[BS_TOK][v.b], step = s.bst[BS_STEP][v.b], gen_step = s.bst[BS_GEN][v.b];
      const float* urow = me.uniforms ? me.uniforms + (size_t)(step + 1) * 16 : nullptr;
      if (mine && tid == 0) {
        s.codes[0] = token;
        s.seen[token >> 5] |= 1u << (token & 31);
      }
      csync();
      // ================= predictor: 15 passes (predictor_graph.py:115-167) =================
      if (P.has_mtp) {
        // pass-0 input rows -> PINB (model dtype): column b = past_hidden, column B + b = codec_embedding(token)
        if (mine && v.rank < 2) {
          for (int t = 0; t < 2; ++t) {
            if (v.gsz >= 2 && t != v.rank) continue;
            uint8_t* dst = reinterpret_cast<uint8_t*>(P.PINB) + (size_t)(t * B + v.b) * HMAX * esz;
            for (int k = tid; k < Ht; k += NCT)
              stw<BF>(dst, k, t == 0 ? __ldcg(me.past_hidden + k) : ldw<BF>(P.t_embed, (size_t)token * Ht + k));
          }
        }
        grid_sync_p(c, PC_GEMV);
        EpiB e{EP_BIAS, P.XB, nullptr, P.ldX, nullptr, 0, P.mtp_b};
        gemv_b<BF>(c, P.seg_mtp, Ht, P.PINB, HMAX, 2 * B, e);
        grid_sync_p(c, PC_NORM);
      }
      for (int i = 0; i < P.ncb; ++i) {
        const int nt = (i == 0) ? 2 : 1;
        // ---- layer-0 input norm of this pass (raw rows -> XB, normalised rows -> XNB)
        pmark(c, PC_NORM);
        if (mine && v.rank < npp) {
          for (int t = 0; t < nt; ++t) {
            const int col = t * B + v.b;
            uint8_t* xn = reinterpret_cast<uint8_t*>(P.XNB) + (size_t)col * P.ldX * esz;
            float* xr = P.XB + (size_t)col * P.ldX;
            if (i == 0 && P.has_mtp) {
              norm_row_b<BF>(c, [&](int k) { return __ldcg(xr + k); }, Sp.ln_in, 0, Sp.H, Sp.eps, xn, nullptr, nullptr, v.rank, npp);
            } else if (i == 0) {
              norm_row_b<BF>(c, [&](int k) { return t == 0 ? __ldcg(me.past_hidden + k) : ldw<BF>(P.t_embed, (size_t)token * Ht + k); },
                             Sp.ln_in, 0, Sp.H, Sp.eps, xn, xr, nullptr, v.rank, npp);
            } else {
              const int prev = s.codes[i];  // code sampled by pass i-1
              const void* tab = P.has_mtp ? P.mtp_tab : P.p_embeds;
              const size_t off = ((size_t)(i - 1) * Sp.V + prev) * (P.has_mtp ? Sp.H : Ht);
              norm_row_b<BF>(c, [&](int k) { return ldw<BF>(tab, off + k); }, Sp.ln_in, 0, Sp.H, Sp.eps, xn, xr, nullptr, v.rank, npp);
            }
          }
        }
        const int slot0 = (i == 0) ? 0 : i + 1;
        stack_b<BF, false>(c, Sp, v, nt, run, slot0);
        {
          EpiB e{EP_F32, P.LOGB, nullptr, VMAX, nullptr, 0, nullptr};
          gemv_b<BF>(c, Sp.seg_head + i, Sp.H, P.XNB, P.ldX, B, e);
        }
        grid_sync_p(c, PC_SAMPLE);
        if (mine) {
          SampleArgs sa;
          sa.logits = P.LOGB + (size_t)v.b * VMAX; sa.V = Sp.V; sa.sp = me.sp_p;
          sa.u = (me.sp_p.do_sample && urow) ? __ldg(urow + 1 + i) : 0.f;
          sa.use_penalty = false; sa.sup0 = Sp.V; sa.suppress_eos = false; sa.eos = -1;
          const int tok = sample_block<BF>(c, sa);
          if (tid == 0) s.codes[i + 1] = tok;
          csync();
        }
      }
      pmark(c, PC_OTHER);
      // ---- emit the frame (generate.py:159): cat(cb0, 15 ids)
      if (mine && v.rank == 0 && tid < 16) me.codes_out[(size_t)s.bst[BS_EMIT][v.b] * 16 + tid] = (long long)s.codes[tid];
      csync();
      // ---- bookkeeping + max_seq_len rule (generate.py:175-177: the frame is already emitted)
      if (warp == 0) {
        bool r = false;
        if (lane < B && ((run >> lane) & 1u)) {
          s.bst[BS_EMIT][lane] += 1;
          const int pos = P.sl[lane].prefill_len + s.bst[BS_STEP][lane];
          if (pos >= P.max_seq_len - 1) {
            s.bst[BS_FIN][lane] = 3;
            s.bst[BS_STEP][lane] += 1;
          } else {
            r = true;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, r);
        if (lane == 0) s.ibc[0] = (int)m;
      }
      csync();
      const uint32_t run2 = (uint32_t)s.ibc[0];
      if (tid == 0) {
        int n = 0;
        for (int b = 0; b < B; ++b)
          if ((run2 >> b) & 1u) s.runl[n++] = b;
        s.runl[32] = n;
      }
      csync();
      if (run2 == 0u) break;
      const bool mine2 = (run2 >> v.b) & 1u;
      // ================= talker step =================
      // layer-0 input: sum of 16 embedding rows + trailing text / tts_pad (generate.py:163-171)
      pmark(c, PC_NORM);
      if (mine2 && v.rank < npt) {
        const void* extra = gen_step < me.trailing_len ? me.trailing : me.tts_pad;
        const size_t eoff = gen_step < me.trailing_len ? (size_t)gen_step * Ht : 0;
        uint8_t* xn = reinterpret_cast<uint8_t*>(P.XNB) + (size_t)v.b * P.ldX * esz;
        float* xr = P.XB + (size_t)v.b * P.ldX;
        norm_row_b<BF>(c, [&](int k) {
          float sm = ldw<BF>(P.t_embed, (size_t)token * Ht + k);
          for (int q = 0; q < P.ncb; ++q) sm += ldw<BF>(P.p_embeds, ((size_t)q * P.p.V + s.codes[q + 1]) * Ht + k);
          return rnd<BF>(rnd<BF>(sm) + ldw<BF>(extra, eoff + k));
        }, P.t.ln_in, 0, Ht, P.t.eps, xn, xr, nullptr, v.rank, npt);
      }
      stack_b<BF, true>(c, P.t, v, 1, run2, 0);
      {
        EpiB e{EP_F32, P.LOGB, nullptr, VMAX, nullptr, 0, nullptr};
        gemv_b<BF>(c, P.t.seg_head, Ht, P.XNB, P.ldX, B, e);
      }
      grid_sync_p(c, PC_SAMPLE);
      if (mine2) {
        SampleArgs sa;
        sa.logits = P.LOGB + (size_t)v.b * VMAX; sa.V = P.t.V; sa.sp = me.sp_t;
        sa.u = (me.sp_t.do_sample && urow) ? __ldg(urow) : 0.f;
        sa.use_penalty = true; sa.sup0 = P.t.V > 1024 ? P.t.V - 1024 : 0;
        sa.suppress_eos = (step + 1) < me.min_new; sa.eos = P.eos;
        const int tok = sample_block<BF>(c, sa);
        if (v.rank == 0 && tid == 0) P.TOKB[v.b] = tok;
      }
      grid_sync_p(c, PC_OTHER);
      if (tid < B && ((run2 >> tid) & 1u)) {
        s.bst[BS_TOK][tid] = __ldcg(P.TOKB + tid);
        s.bst[BS_STEP][tid] += 1;
        s.bst[BS_GEN][tid] += 1;
      }
      csync();
    }
    pmark(c, PC_OTHER);
    if ((P.dbg_on & 2) && cta == 0 && tid == 0 && P.mode != MODE_GEMV_TEST)
      for (int i = 0; i < PC_N; ++i) reinterpret_cast<long long*>(P.dbg)[i] += s.prof[2 + i];   // accumulates over launches
    if (v.rank == 0 && P.mode != MODE_GEMV_TEST) {
      if (tid == 0) {
        int* st = me.state;
        st[0] = s.bst[BS_TOK][v.b]; st[1] = s.bst[BS_STEP][v.b]; st[2] = s.bst[BS_GEN][v.b];
        st[3] = s.bst[BS_FIN][v.b]; st[4] = s.bst[BS_EMIT][v.b];
      }
      for (int i = tid; i < VMAX / 32; i += NCT) me.seen[i] = s.seen[i];
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
