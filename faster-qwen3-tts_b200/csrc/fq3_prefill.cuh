// fq3_prefill.cuh -- K3: hand-written talker prefill (bf16).  Included at the end of fq3_engine.cu.
//
// Replaces the variable-length prompt forward the reference delegates to upstream HF eager code
// (faster_qwen3_tts/generate.py:107-118, streaming.py:63-74) and the 56 index_copy_ calls of
// TalkerGraph.prefill_kv (talker_graph.py:153-170): K and V are written straight into the engine's
// [layer][kv_head][slot][128] cache.  Per layer: RMSNorm rows -> QKV GEMM -> q/k-norm + RoPE + KV append ->
// causal GQA attention (eager semantics: bf16 scores, fp32 softmax rounded to bf16, bf16 P.V) -> o_proj GEMM with
// fused residual -> RMSNorm rows -> gate/up GEMM with fused SwiGLU (interleaved columns) -> down GEMM with fused
// residual.  GEMMs are the shared implicit-GEMM tensor-core kernel (fq3_gemm.cuh, taps = 1).
#pragma once
#include "fq3_gemm.cuh"
#include "fq3_gemm_tc.cuh"

namespace pf {

__device__ __forceinline__ float rb(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// one block (256 threads) per row: Y = w * rnd(x * rsqrt(mean(x^2) + eps))
__global__ void rmsnorm_rows_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ w, int H,
                                    float eps, __nv_bfloat16* __restrict__ Y) {
  __shared__ float red[8];
  const size_t row = blockIdx.x;
  const int tid = threadIdx.x;
  float v[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = tid + i * 256;
    v[i] = k < H ? __bfloat162float(X[row * H + k]) : 0.f;
    ss += v[i] * v[i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float r = 1.0f / sqrtf(tot / (float)H + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = tid + i * 256;
    if (k < H) Y[row * H + k] = __float2bfloat16_rn(__bfloat162float(w[k]) * rb(v[i] * r));
  }
}

// one warp per (token, vector) with vector in [q heads | k heads | v heads]: q/k RMSNorm + RoPE in place (q) or into
// the KV cache (k, v).  QKV is [P][qd + 2 kd] bf16.
__global__ void rope_kv_kernel(__nv_bfloat16* __restrict__ QKV, int P, int nH, int nKV, const __nv_bfloat16* qn,
                               const __nv_bfloat16* kn, const float* __restrict__ cosT, const float* __restrict__ sinT,
                               int npos, int n_left_pad, float eps, __nv_bfloat16* __restrict__ kc,
                               __nv_bfloat16* __restrict__ vc, int S) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nvec = nH + 2 * nKV;
  if (gw >= P * nvec) return;
  const int t = gw / nvec, vi = gw % nvec;
  const int ld = (nH + 2 * nKV) * 128;
  __nv_bfloat16* src = QKV + (size_t)t * ld + (size_t)vi * 128;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __bfloat162float(src[lane + 32 * i]);
  const int what = vi < nH ? 0 : (vi < nH + nKV ? 1 : 2);
  if (what < 2) {
    float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float r = 1.0f / sqrtf(ss / 128.0f + eps);
    const __nv_bfloat16* nw = what == 0 ? qn : kn;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = rb(__bfloat162float(nw[lane + 32 * i]) * rb(v[i] * r));
    int rp = t - n_left_pad;
    rp = rp < 0 ? 0 : (rp >= npos ? npos - 1 : rp);
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      const float cc = rb(cosT[(size_t)rp * 128 + e]), sv = rb(sinT[(size_t)rp * 128 + e]);
      const float rot = (i < 2) ? -v[i + 2] : v[i - 2];
      o[i] = rb(rb(v[i] * cc) + rb(rot * sv));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = o[i];
  }
  if (what == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) src[lane + 32 * i] = __float2bfloat16_rn(v[i]);
  } else {
    const int g = what == 1 ? vi - nH : vi - nH - nKV;
    __nv_bfloat16* dst = (what == 1 ? kc : vc) + ((size_t)g * S + t) * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = __float2bfloat16_rn(v[i]);
  }
}

// causal GQA attention over the cache; block = (8 queries) x (1 head), warp w owns query i0 + w.
// dynamic smem: scores [8][Ppad] fp32 + q [8][128] fp32
__global__ void __launch_bounds__(256) attn_prefill_kernel(const __nv_bfloat16* __restrict__ QKV, int P, int nH, int nKV,
                                                          const __nv_bfloat16* __restrict__ kc,
                                                          const __nv_bfloat16* __restrict__ vc, int S, int n_left_pad,
                                                          __nv_bfloat16* __restrict__ OUT) {
  extern __shared__ float sm[];
  const int Ppad = (P + 31) & ~31;
  float* sc = sm;                 // [8][Ppad]
  float* qs = sm + 8 * Ppad;      // [8][128]
  const int h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  const int g = h / (nH / nKV);
  const int ld = (nH + 2 * nKV) * 128;
  const bool active = i < P;
  if (active)
#pragma unroll
    for (int e = 0; e < 4; ++e) qs[warp * 128 + lane + 32 * e] = __bfloat162float(QKV[(size_t)i * ld + h * 128 + lane + 32 * e]);
  __syncwarp();
  if (!active) return;
  const __nv_bfloat16* kb = kc + (size_t)g * S * 128;
  const __nv_bfloat16* vb = vc + (size_t)g * S * 128;
  const float scale = 0.08838834764831845f;
  float* my = sc + warp * Ppad;
  const float* q = qs + warp * 128;
  // scores: lane handles keys j = n_left_pad + lane, +32, ...
  float mx = -INFINITY;
  for (int j = n_left_pad + lane; j <= i; j += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(kb + (size_t)j * 128);
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint4 w = __ldg(kr + c);
      const float4 qa = *reinterpret_cast<const float4*>(q + c * 8), qb = *reinterpret_cast<const float4*>(q + c * 8 + 4);
      d = fmaf(qa.x, __uint_as_float(w.x << 16), d); d = fmaf(qa.y, __uint_as_float(w.x & 0xffff0000u), d);
      d = fmaf(qa.z, __uint_as_float(w.y << 16), d); d = fmaf(qa.w, __uint_as_float(w.y & 0xffff0000u), d);
      d = fmaf(qb.x, __uint_as_float(w.z << 16), d); d = fmaf(qb.y, __uint_as_float(w.z & 0xffff0000u), d);
      d = fmaf(qb.z, __uint_as_float(w.w << 16), d); d = fmaf(qb.w, __uint_as_float(w.w & 0xffff0000u), d);
    }
    const float s = rb(rb(d) * scale);
    my[j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = n_left_pad + lane; j <= i; j += 32) {
    const float e = expf(my[j] - mx);
    my[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  // P.V: lane owns dims [4*lane, 4*lane+4)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int j = n_left_pad; j <= i; ++j) {
    const float p = rb(my[j] / sum);
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(vb + (size_t)j * 128) + lane);
    a0 = fmaf(p, __uint_as_float(w.x << 16), a0); a1 = fmaf(p, __uint_as_float(w.x & 0xffff0000u), a1);
    a2 = fmaf(p, __uint_as_float(w.y << 16), a2); a3 = fmaf(p, __uint_as_float(w.y & 0xffff0000u), a3);
  }
  __nv_bfloat16* o = OUT + (size_t)i * nH * 128 + h * 128 + 4 * lane;
  o[0] = __float2bfloat16_rn(a0); o[1] = __float2bfloat16_rn(a1); o[2] = __float2bfloat16_rn(a2); o[3] = __float2bfloat16_rn(a3);
}

}  // namespace pf

static int pf_gemm(fq3_engine* e, const __nv_bfloat16* X, const __nv_bfloat16* W, const __nv_bfloat16* R,
                   __nv_bfloat16* Y, int T, int K, int N, int mode, cudaStream_t stream) {
  fq3gemm::ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.X = X; a.W = W; a.R = R; a.Yraw = Y; a.T = T; a.Cin = K; a.N = N; a.taps = 1; a.dil = 1;
  a.bias_mod = 1; a.act_mod = 1; a.mode = mode;
  e->launches++;
  if (g_fq3_gemm_backend != 1) {
    const int r = fq3tc::launch_tc(a, stream, g_fq3_gemm_backend);
    if (r == 0) return 0;
    if (r < 0) return fail(FQ3_ERR_CUDA, "tcgen05 GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  dim3 grid((T + fq3gemm::BM - 1) / fq3gemm::BM, (N + fq3gemm::BN - 1) / fq3gemm::BN);
  fq3gemm::conv_gemm_kernel<<<grid, fq3gemm::CTHREADS, fq3gemm::CONV_SMEM, stream>>>(a);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int fq3_engine_set_prefill_weights(fq3_engine* e, const fq3_tensor* tensors, int32_t n) {
  if (!e || !tensors) return fail(FQ3_ERR_INVALID, "null argument");
  if (!e->bf16) return fail(FQ3_ERR_INVALID, "the hand-written prefill is bf16 only");
  const fq3_stack_config& T = e->cfg.talker;
  const int64_t L = T.num_hidden_layers, H = T.hidden_size, I = T.intermediate_size;
  const int64_t qd = T.num_attention_heads * 128, kd = T.num_key_value_heads * 128;
  struct Want { const char* nm; int64_t numel; const void** dst; } want[] = {
      {"t.qkv", L * (qd + 2 * kd) * H, &e->pf_qkv}, {"t.o", L * H * qd, &e->pf_o}, {"t.gu", L * 2 * I * H, &e->pf_gu},
      {"t.down", L * H * I, &e->pf_down}, {"t.head", (int64_t)T.vocab_size * H, &e->pf_head}};
  for (auto& w : want) {
    *w.dst = nullptr;
    for (int i = 0; i < n; ++i)
      if (!strcmp(tensors[i].name, w.nm)) {
        if (tensors[i].numel != w.numel) return fail(FQ3_ERR_INVALID, "prefill tensor '%s': bad numel", w.nm);
        *w.dst = tensors[i].dev_ptr;
      }
    if (!*w.dst) return fail(FQ3_ERR_INVALID, "missing prefill tensor '%s'", w.nm);
  }
  if (H > 2048 || H % 32 || I % 32) return fail(FQ3_ERR_INVALID, "prefill geometry unsupported");
  DevGuard dev_guard(e->dev);
  CK(cudaFuncSetAttribute(fq3gemm::conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fq3gemm::CONV_SMEM));
  const size_t S = e->cfg.max_seq_len;
  const size_t wide = std::max<size_t>(qd + 2 * kd, (size_t)I);
  if (!e->pf_buf[0]) {
    CK(cudaMalloc(&e->pf_buf[0], S * H * 2));      // x
    CK(cudaMalloc(&e->pf_buf[1], S * H * 2));      // x1
    CK(cudaMalloc(&e->pf_buf[2], S * H * 2));      // normed
    CK(cudaMalloc(&e->pf_buf[3], S * wide * 2));   // qkv / act
    CK(cudaMalloc(&e->pf_buf[4], S * qd * 2));     // attention out
  }
  // per FUNCTION, not per engine: size it for the largest cache any engine may have (SEQMAX), so a second engine
  // with a shorter max_seq_len cannot lower the limit under the first one
  CK(cudaFuncSetAttribute(pf::attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)((8 * (size_t)fq3::SEQMAX + 8 * 128) * sizeof(float))));
  e->pf_ready = true;
  return 0;
}

extern "C" int fq3_prefill(fq3_engine* e, int32_t slot, const void* embeds_dev, int32_t P, int32_t n_left_pad,
                           void* logits_out_dev, void* hidden_out_dev, void* stream_) {
  if (!e || !embeds_dev || !logits_out_dev || !hidden_out_dev) return fail(FQ3_ERR_INVALID, "null argument");
  {
    int rcs;
    if ((rcs = check_slot(e, slot))) return rcs;
  }
  if (!e->pf_ready) return fail(FQ3_ERR_STATE, "fq3_engine_set_prefill_weights has not been called");
  if (P <= 0) return fail(FQ3_ERR_INVALID, "empty prompt");
  if (P > e->cfg.max_seq_len)
    return fail(FQ3_ERR_TOO_LONG, "Input is too long: prefill has %d tokens but max_seq_len=%d. Use shorter text or shorter reference audio.", P, e->cfg.max_seq_len);
  DevGuard dev_guard(e->dev);
  cudaStream_t stream = (cudaStream_t)stream_;
  const fq3_stack_config& T = e->cfg.talker;
  const int L = T.num_hidden_layers, H = T.hidden_size, I = T.intermediate_size;
  const int nH = T.num_attention_heads, nKV = T.num_key_value_heads, qd = nH * 128, kd = nKV * 128, S = e->cfg.max_seq_len;
  using bf = __nv_bfloat16;
  bf* x = (bf*)e->pf_buf[0];
  bf* x1 = (bf*)e->pf_buf[1];
  bf* hn = (bf*)e->pf_buf[2];
  bf* wide = (bf*)e->pf_buf[3];
  bf* att = (bf*)e->pf_buf[4];
  CK(cudaMemcpyAsync(x, embeds_dev, (size_t)P * H * 2, cudaMemcpyDeviceToDevice, stream));
  const KParams& k = e->kp;
  const int Ppad = (P + 31) & ~31;
  const size_t attn_smem = (size_t)(8 * Ppad + 8 * 128) * sizeof(float);
  int rc;
  for (int l = 0; l < L; ++l) {
    pf::rmsnorm_rows_kernel<<<P, 256, 0, stream>>>(x, (const bf*)k.t.ln_in + (size_t)l * H, H, T.rms_norm_eps, hn);
    e->launches++;
    if ((rc = pf_gemm(e, hn, (const bf*)e->pf_qkv + (size_t)l * (qd + 2 * kd) * H, nullptr, wide, P, H, qd + 2 * kd, 0, stream))) return rc;
    {
      const int warps = P * (nH + 2 * nKV);
      pf::rope_kv_kernel<<<(warps * 32 + 255) / 256, 256, 0, stream>>>(
          wide, P, nH, nKV, (const bf*)k.t.qnorm + (size_t)l * 128, (const bf*)k.t.knorm + (size_t)l * 128, k.t.cos,
          k.t.sin, k.t.npos, n_left_pad, T.rms_norm_eps, (bf*)slot_tk(e, slot) + (size_t)l * nKV * S * 128,
          (bf*)slot_tv(e, slot) + (size_t)l * nKV * S * 128, S);
      e->launches++;
    }
    pf::attn_prefill_kernel<<<dim3((P + 7) / 8, nH), 256, attn_smem, stream>>>(
        wide, P, nH, nKV, (const bf*)slot_tk(e, slot) + (size_t)l * nKV * S * 128, (const bf*)slot_tv(e, slot) + (size_t)l * nKV * S * 128, S,
        n_left_pad, att);
    e->launches++;
    if ((rc = pf_gemm(e, att, (const bf*)e->pf_o + (size_t)l * H * qd, x, x1, P, qd, H, 0, stream))) return rc;
    pf::rmsnorm_rows_kernel<<<P, 256, 0, stream>>>(x1, (const bf*)k.t.ln_post + (size_t)l * H, H, T.rms_norm_eps, hn);
    e->launches++;
    if ((rc = pf_gemm(e, hn, (const bf*)e->pf_gu + (size_t)l * 2 * I * H, nullptr, wide, P, H, 2 * I, 1, stream))) return rc;
    if ((rc = pf_gemm(e, wide, (const bf*)e->pf_down + (size_t)l * H * I, x1, x, P, I, H, 0, stream))) return rc;
  }
  // final norm of the last row -> past_hidden; logits = codec_head(hidden)
  pf::rmsnorm_rows_kernel<<<1, 256, 0, stream>>>(x + (size_t)(P - 1) * H, (const bf*)k.t.ln_f, H, T.rms_norm_eps, hn);
  e->launches++;
  CK(cudaMemcpyAsync(hidden_out_dev, hn, (size_t)H * 2, cudaMemcpyDeviceToDevice, stream));
  if ((rc = pf_gemm(e, hn, (const bf*)e->pf_head, nullptr, (bf*)logits_out_dev, 1, H, T.vocab_size, 0, stream))) return rc;
  CK(cudaGetLastError());
  return 0;
}
