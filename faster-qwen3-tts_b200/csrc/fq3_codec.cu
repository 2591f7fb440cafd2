// fq3_codec.cu -- K4: the waveform decoder stack of the codec (conv_in -> 4 x [SnakeBeta, causal ConvTranspose,
// 3 residual units] -> SnakeBeta -> conv_out -> clamp) as hand-written sm_100a kernels behind the C ABI.
// Replaces the cuDNN/cuBLAS launches issued by the reference's `speech_tokenizer.decode` call sites
// (faster_qwen3_tts/model.py:924,1093,1122) for the FLOP-dominant part of the decoder (94% of its FLOPs).
//
// One kernel does every dense layer: a causal conv1d as an implicit GEMM over channels-last bf16 activations
//     Y[t, n] = bias[n] + sum_{tap, ci} W[n, tap, ci] * X[t - (taps-1-tap)*dil, ci]        (X[<0] = 0)
//   * M = time, N = output channels, K = taps x Cin; 128 x 96 x 32 tiles, 8 warps (2 x 4), bf16 mma.sync
//     m16n8k16 with fp32 accumulation, ldmatrix from XOR-swizzled shared memory, 4-stage cp.async pipeline.
//   * A causal ConvTranspose1d(k = 2r, stride r) is the same kernel with 2 taps and N' = r*Cout "phase" channels;
//     the [T, r*Cout] result IS the [T*r, Cout] upsampled sequence (pixel shuffle is a reinterpretation).
//   * Epilogue fuses bias, residual add, and the NEXT layer's SnakeBeta (x + sin^2(a x) / (b + eps)), writing the raw
//     and/or the activated tensor, so no element-wise kernel exists in the stack.
// conv_out (96 -> 1 channel) + clamp is a small dedicated kernel.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/fq3_engine.h"
#include "fq3_gemm.cuh"
#include "fq3_gemm_tc.cuh"

extern int g_fq3_gemm_backend;  // 0 = tcgen05/TMA kernel when the shape allows, 1 = force the mma.sync kernel

namespace {

thread_local char g_cerr[512] = "";
int cfail(int code, const char* msg, const char* extra = "") {
  snprintf(g_cerr, sizeof(g_cerr), "%s%s", msg, extra);
  return code;
}
#define CCK(call)                                                                  \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return cfail(FQ3_ERR_CUDA, #call " failed: ", cudaGetErrorString(_e)); \
  } while (0)

using namespace fq3gemm;

// final causal conv7 (C -> 1) over activated input + clamp to [-1, 1]; one thread per output sample
__global__ void conv_out_kernel(const __nv_bfloat16* __restrict__ X, const float* __restrict__ W, float bias, int T,
                                int C, int taps, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  X += (size_t)blockIdx.y * T * C;      // blockIdx.y = sequence of the batch
  out += (size_t)blockIdx.y * T;
  float s = bias;
  for (int k = 0; k < taps; ++k) {
    const int tt = t - (taps - 1 - k);
    if (tt < 0) continue;
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(X + (size_t)tt * C);
    const float* w = W + (size_t)k * C;
    for (int c = 0; c < C / 2; ++c) {
      const __nv_bfloat162 v = x[c];
      s = fmaf(__bfloat162float(v.x), w[2 * c], s);
      s = fmaf(__bfloat162float(v.y), w[2 * c + 1], s);
    }
  }
  out[t] = fminf(1.f, fmaxf(-1.f, s));
}

// first SnakeBeta applied to the bf16 input of the stack is folded into conv_in's epilogue; the stack input itself
// (output of the upsampling front end) arrives channels-first from torch -> transpose + cast here
__global__ void to_channels_last_kernel(const __nv_bfloat16* __restrict__ X, int C, int T, __nv_bfloat16* __restrict__ Y) {
  __shared__ __nv_bfloat16 tile[32][33];
  X += (size_t)blockIdx.z * C * T;      // blockIdx.z = sequence of the batch
  Y += (size_t)blockIdx.z * C * T;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? X[(size_t)c * T + t] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) Y[(size_t)t * C + c] = tile[threadIdx.x][i];
  }
}

struct Layer {
  int Cin, N, taps, dil, bias_mod, act_mod;  // act_mod 0 => no activated output
  bool write_raw, residual;
  int upsample;                              // r for ConvTranspose-as-conv (T_out = T*r, channels N/r), else 1
  __nv_bfloat16* W = nullptr;
  float *bias = nullptr, *ea = nullptr, *ib = nullptr;
};

}  // namespace

struct fq3_codec {
  int dev = 0;
  int hidden = 0, decoder_dim = 0, n_blocks = 0;
  int rates[8];
  std::vector<Layer> layers;
  float* w_out = nullptr;
  float b_out = 0.f;
  int c_out = 0;
  // scratch (grown on demand)
  size_t cap = 0;
  __nv_bfloat16 *buf[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t launches = 0;
  double flops_per_frame = 0;
  std::vector<void*> owned;
};

extern "C" int fq3_codec_create(const int32_t* geom, int32_t n_geom, fq3_codec** out) {
  // geom = {device, hidden_size, decoder_dim, n_blocks, rate_0 .. rate_{n-1}}
  if (!geom || !out || n_geom < 5) return cfail(FQ3_ERR_INVALID, "bad codec geometry");
  fq3_codec* c = new fq3_codec();
  c->dev = geom[0]; c->hidden = geom[1]; c->decoder_dim = geom[2]; c->n_blocks = geom[3];
  if (c->n_blocks < 1 || c->n_blocks > 8 || n_geom < 4 + c->n_blocks) { delete c; return cfail(FQ3_ERR_INVALID, "bad codec geometry"); }
  for (int i = 0; i < c->n_blocks; ++i) c->rates[i] = geom[4 + i];
  if (c->hidden % BK || c->decoder_dim % (BK << c->n_blocks)) { delete c; return cfail(FQ3_ERR_INVALID, "codec channels must be multiples of 32 at every level"); }
  CCK(cudaSetDevice(c->dev));
  CCK(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
  *out = c;
  return 0;
}

extern "C" void fq3_codec_destroy(fq3_codec* c) {
  if (!c) return;
  cudaSetDevice(c->dev);
  for (void* p : c->owned) cudaFree(p);
  for (auto* b : c->buf) if (b) cudaFree(b);
  delete c;
}

// tensors (all float32 on device, PyTorch layouts; the engine converts / rearranges):
//   conv_in.w [D,H,7] conv_in.b [D]
//   b{i}.act.a b{i}.act.b [Cin]      b{i}.up.w [Cin,Cout,2r] b{i}.up.b [Cout]
//   b{i}.r{j}.a1.a .a1.b [C]  .c1.w [C,C,7] .c1.b [C]  .a2.a .a2.b [C]  .c2.w [C,C,1] .c2.b [C]
//   out.act.a out.act.b [C]   out.w [1,C,7] out.b [1]
extern "C" int fq3_codec_load_weights(fq3_codec* c, const fq3_tensor* tensors, int32_t n, void* stream_) {
  if (!c || !tensors) return cfail(FQ3_ERR_INVALID, "null argument");
  CCK(cudaSetDevice(c->dev));
  cudaStream_t stream = (cudaStream_t)stream_;
  std::map<std::string, const fq3_tensor*> tm;
  for (int i = 0; i < n; ++i) tm[tensors[i].name] = &tensors[i];
  std::vector<std::vector<float>> keep;
  auto host = [&](const std::string& nm, int64_t numel, std::vector<float>& dst) -> int {
    auto it = tm.find(nm);
    if (it == tm.end()) return cfail(FQ3_ERR_INVALID, "missing codec tensor ", nm.c_str());
    if (it->second->numel != numel) return cfail(FQ3_ERR_INVALID, "bad numel for codec tensor ", nm.c_str());
    dst.resize(numel);
    CCK(cudaMemcpyAsync(dst.data(), it->second->dev_ptr, numel * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CCK(cudaStreamSynchronize(stream));
    return 0;
  };
  auto up_f = [&](const std::vector<float>& v, float** d) -> int {
    CCK(cudaMalloc(d, v.size() * sizeof(float)));
    c->owned.push_back(*d);
    CCK(cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    CCK(cudaStreamSynchronize(stream));
    return 0;
  };
  auto up_bf = [&](const std::vector<float>& v, __nv_bfloat16** d) -> int {
    std::vector<__nv_bfloat16> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16(v[i]);
    CCK(cudaMalloc(d, h.size() * 2));
    c->owned.push_back(*d);
    CCK(cudaMemcpyAsync(*d, h.data(), h.size() * 2, cudaMemcpyHostToDevice, stream));
    CCK(cudaStreamSynchronize(stream));
    return 0;
  };
  auto snake = [&](const std::string& pre, int C, float** ea, float** ib) -> int {
    std::vector<float> al, be;
    int rc;
    if ((rc = host(pre + ".a", C, al))) return rc;
    if ((rc = host(pre + ".b", C, be))) return rc;
    for (int i = 0; i < C; ++i) { al[i] = expf(al[i]); be[i] = 1.0f / (expf(be[i]) + 1e-9f); }
    if ((rc = up_f(al, ea))) return rc;
    return up_f(be, ib);
  };
  // causal conv weight [N, Cin, taps] (PyTorch) -> [N][taps][Cin]
  auto conv_w = [&](const std::string& nm, int N, int Cin, int taps, __nv_bfloat16** d) -> int {
    std::vector<float> w, r((size_t)N * taps * Cin);
    int rc;
    if ((rc = host(nm, (int64_t)N * Cin * taps, w))) return rc;
    for (int nn = 0; nn < N; ++nn)
      for (int ci = 0; ci < Cin; ++ci)
        for (int k = 0; k < taps; ++k) r[((size_t)nn * taps + k) * Cin + ci] = w[((size_t)nn * Cin + ci) * taps + k];
    return up_bf(r, d);
  };
  c->layers.clear();
  int rc;
  const int H = c->hidden, D = c->decoder_dim;
  double flops = 0;  // per input frame position of the stack (multiply by 4*T)
  double pos = 1.0;  // positions per stack-input position
  {  // conv_in: writes only the block-0 SnakeBeta-activated tensor
    Layer L{H, D, 7, 1, D, D, false, false, 1};
    if ((rc = conv_w("conv_in.w", D, H, 7, &L.W))) return rc;
    std::vector<float> b;
    if ((rc = host("conv_in.b", D, b))) return rc;
    if ((rc = up_f(b, &L.bias))) return rc;
    if ((rc = snake("b0.act", D, &L.ea, &L.ib))) return rc;
    c->layers.push_back(L);
    flops += 2.0 * D * H * 7 * pos;
  }
  int Cin = D;
  for (int bi = 0; bi < c->n_blocks; ++bi) {
    const int Cout = Cin / 2, r = c->rates[bi];
    const std::string p = "b" + std::to_string(bi);
    {  // ConvTranspose(k=2r, s=r) as a 2-tap conv with N' = r*Cout
      Layer L{Cin, r * Cout, 2, 1, Cout, Cout, true, false, r};
      std::vector<float> w, wr((size_t)r * Cout * 2 * Cin);
      if ((rc = host(p + ".up.w", (int64_t)Cin * Cout * 2 * r, w))) return rc;
      for (int j = 0; j < r; ++j)
        for (int co = 0; co < Cout; ++co)
          for (int ci = 0; ci < Cin; ++ci) {
            const size_t nrow = (size_t)j * Cout + co;
            wr[(nrow * 2 + 0) * Cin + ci] = w[((size_t)ci * Cout + co) * 2 * r + j + r];  // tap 0 <-> x[q-1]
            wr[(nrow * 2 + 1) * Cin + ci] = w[((size_t)ci * Cout + co) * 2 * r + j];      // tap 1 <-> x[q]
          }
      if ((rc = up_bf(wr, &L.W))) return rc;
      std::vector<float> b;
      if ((rc = host(p + ".up.b", Cout, b))) return rc;
      if ((rc = up_f(b, &L.bias))) return rc;
      if ((rc = snake(p + ".r0.a1", Cout, &L.ea, &L.ib))) return rc;
      c->layers.push_back(L);
      flops += 2.0 * r * Cout * 2 * Cin * pos;
      pos *= r;
    }
    static const int dils[3] = {1, 3, 9};
    for (int j = 0; j < 3; ++j) {
      const std::string q = p + ".r" + std::to_string(j);
      {  // conv7(dil) on act1(x) -> act2 fused
        Layer L{Cout, Cout, 7, dils[j], Cout, Cout, false, false, 1};
        if ((rc = conv_w(q + ".c1.w", Cout, Cout, 7, &L.W))) return rc;
        std::vector<float> b;
        if ((rc = host(q + ".c1.b", Cout, b))) return rc;
        if ((rc = up_f(b, &L.bias))) return rc;
        if ((rc = snake(q + ".a2", Cout, &L.ea, &L.ib))) return rc;
        c->layers.push_back(L);
      }
      {  // conv1 + residual -> raw (next residual) + the next consumer's SnakeBeta
        const bool last = j == 2;
        Layer L{Cout, Cout, 1, 1, Cout, Cout, !last, true, 1};
        if ((rc = conv_w(q + ".c2.w", Cout, Cout, 1, &L.W))) return rc;
        std::vector<float> b;
        if ((rc = host(q + ".c2.b", Cout, b))) return rc;
        if ((rc = up_f(b, &L.bias))) return rc;
        const std::string nxt = !last ? p + ".r" + std::to_string(j + 1) + ".a1"
                                      : (bi + 1 < c->n_blocks ? "b" + std::to_string(bi + 1) + ".act" : std::string("out.act"));
        if ((rc = snake(nxt, Cout, &L.ea, &L.ib))) return rc;
        c->layers.push_back(L);
      }
      flops += 2.0 * Cout * Cout * 8 * pos;
    }
    Cin = Cout;
  }
  {
    std::vector<float> w, wr((size_t)7 * Cin), b;
    if ((rc = host("out.w", (int64_t)Cin * 7, w))) return rc;
    for (int ci = 0; ci < Cin; ++ci)
      for (int k = 0; k < 7; ++k) wr[(size_t)k * Cin + ci] = w[(size_t)ci * 7 + k];
    if ((rc = up_f(wr, &c->w_out))) return rc;
    if ((rc = host("out.b", 1, b))) return rc;
    c->b_out = b[0];
    c->c_out = Cin;
    flops += 2.0 * Cin * 7 * pos;
  }
  c->flops_per_frame = flops;
  return 0;
}

static int launch_conv(fq3_codec* c, const Layer& L, const __nv_bfloat16* X, const __nv_bfloat16* R, __nv_bfloat16* Yraw,
                       __nv_bfloat16* Yact, int T, int batch, cudaStream_t stream) {
  ConvArgs a;
  a.X = X; a.W = L.W; a.bias = L.bias; a.R = R; a.Yraw = Yraw; a.Yact = Yact; a.ea = L.ea; a.ib = L.ib;
  a.T = T; a.Cin = L.Cin; a.N = L.N; a.taps = L.taps; a.dil = L.dil; a.bias_mod = L.bias_mod; a.act_mod = L.act_mod;
  a.mode = 0;
  a.batch = batch;
  c->launches++;
  if (g_fq3_gemm_backend == 0) {
    const int r = fq3tc::launch_tc(a, stream);
    if (r == 0) return 0;
    if (r < 0) return cfail(FQ3_ERR_CUDA, "tcgen05 conv launch failed: ", cudaGetErrorString(cudaGetLastError()));
  }
  dim3 grid(((T + BM - 1) / BM) * batch, (L.N + BN - 1) / BN);
  conv_gemm_kernel<<<grid, CTHREADS, CONV_SMEM, stream>>>(a);
  CCK(cudaGetLastError());
  return 0;
}

// x_dev: [batch][hidden][T4] bf16 channels-first (output of the front end), pcm_out_dev float32 [batch][T4 * prod(rates)]:
// `batch` independent windows of equal length share every launch (each with its own causal left padding)
extern "C" int fq3_codec_decode_batch(fq3_codec* c, const void* x_dev, int32_t batch, int32_t T4, float* pcm_out_dev,
                                      void* stream_) {
  if (!c || !x_dev || !pcm_out_dev || T4 <= 0 || batch <= 0) return cfail(FQ3_ERR_INVALID, "null argument");
  if (c->layers.empty()) return cfail(FQ3_ERR_STATE, "codec weights not loaded");
  CCK(cudaSetDevice(c->dev));
  cudaStream_t stream = (cudaStream_t)stream_;
  // largest tensor: [T_final][C_final * 2] worth of bf16 at the widest level; size every buffer for the maximum
  size_t need = (size_t)T4 * c->hidden;
  {
    size_t T = T4;
    int C = c->decoder_dim;
    need = std::max(need, T * (size_t)C);
    for (int bi = 0; bi < c->n_blocks; ++bi) {
      T *= c->rates[bi];
      C /= 2;
      need = std::max(need, T * (size_t)C);
    }
  }
  need *= (size_t)batch;
  if (need > c->cap) {
    for (auto*& b : c->buf) { if (b) cudaFree(b); b = nullptr; }
    for (auto*& b : c->buf) CCK(cudaMalloc(&b, need * 2));
    c->cap = need;
  }
  {
    dim3 g((T4 + 31) / 32, (c->hidden + 31) / 32, batch), b(32, 8);
    to_channels_last_kernel<<<g, b, 0, stream>>>((const __nv_bfloat16*)x_dev, c->hidden, T4, c->buf[0]);
    c->launches++;
  }
  int rc;
  int T = T4;
  size_t li = 0;
  // four ping-pong buffers.  `cur` always holds the activated input of the next layer; the other three are free.
  __nv_bfloat16* cur = c->buf[1];
  if ((rc = launch_conv(c, c->layers[li++], c->buf[0], nullptr, nullptr, cur, T, batch, stream))) return rc;
  for (int bi = 0; bi < c->n_blocks; ++bi) {
    __nv_bfloat16* f[3];
    int k = 0;
    for (auto* b : c->buf)
      if (b != cur) f[k++] = b;
    __nv_bfloat16 *x = f[0], *a1 = f[1], *a2 = f[2], *y = cur;  // cur is free once the up-conv has consumed it
    const Layer& U = c->layers[li++];
    if ((rc = launch_conv(c, U, cur, nullptr, x, a1, T, batch, stream))) return rc;  // raw -> x, SnakeBeta(raw) -> a1
    T *= U.upsample;
    for (int j = 0; j < 3; ++j) {
      const Layer& C1 = c->layers[li++];
      const Layer& C2 = c->layers[li++];
      if ((rc = launch_conv(c, C1, a1, nullptr, nullptr, a2, T, batch, stream))) return rc;  // a2 = act2(conv7(a1))
      // y = conv1(a2) + x ; a1 <- SnakeBeta_next(y)   (conv7 has consumed a1, so it can be overwritten)
      if ((rc = launch_conv(c, C2, a2, x, C2.write_raw ? y : nullptr, a1, T, batch, stream))) return rc;
      std::swap(x, y);
    }
    cur = a1;
  }
  __nv_bfloat16* act = cur;
  conv_out_kernel<<<dim3((T + 255) / 256, batch), 256, 0, stream>>>(act, c->w_out, c->b_out, T, c->c_out, 7, pcm_out_dev);
  c->launches++;
  CCK(cudaGetLastError());
  return 0;
}

extern "C" int fq3_codec_decode(fq3_codec* c, const void* x_dev, int32_t T4, float* pcm_out_dev, void* stream_) {
  return fq3_codec_decode_batch(c, x_dev, 1, T4, pcm_out_dev, stream_);
}

extern "C" double fq3_codec_flops(fq3_codec* c, int32_t T4) { return c ? c->flops_per_frame * T4 : 0.0; }
extern "C" int64_t fq3_codec_launch_count(fq3_codec* c) { return c ? c->launches : 0; }
extern "C" const char* fq3_codec_last_error(void) { return g_cerr; }
