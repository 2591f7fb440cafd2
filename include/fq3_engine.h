/*
 * fq3_engine.h -- C ABI of the B200-native Qwen3-TTS decode engine.
 *
 * Drop-in boundary for the hot path of andimarafioti/faster-qwen3-tts.  The reference has no FFI on its torch
 * path (the seam is a Python duck type); the nearest precedent is the qwentts.cpp C ABI it reaches through
 * ctypes at faster_qwen3_tts/ggml_backend.py:216,381,446,499,645.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 on success, a negative fq3_status on failure; fq3_last_error() gives the text.
 *     Nothing throws across the ABI.
 *   - "dev" pointers are device memory owned by the caller (torch); the engine owns its packed weights, KV
 *     caches, scratch and per-request state.  `stream` is a cudaStream_t passed as void* (0 = default stream).
 *   - one engine per device; calls on one engine are not re-entrant; work is stream-ordered, the only host
 *     synchronisation is inside fq3_decode_chunk (it returns its result to host memory).
 *   - model dtype (FQ3_F32 / FQ3_BF16) is fixed at create time; all weight / activation tensors crossing the
 *     ABI are in that dtype, row-major, unless a parameter says otherwise.
 */
#ifndef FQ3_ENGINE_H
#define FQ3_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fq3_engine fq3_engine;

enum fq3_status {
  FQ3_OK = 0,
  FQ3_ERR_INVALID = -1,   /* bad argument / unsupported geometry           */
  FQ3_ERR_CUDA = -2,      /* CUDA runtime error                            */
  FQ3_ERR_STATE = -3,     /* call order violated (e.g. decode before load) */
  FQ3_ERR_TOO_LONG = -4   /* prompt longer than max_seq_len (talker_graph.py:163-167 raises RuntimeError) */
};

enum fq3_dtype { FQ3_F32 = 0, FQ3_BF16 = 1 };

/* One transformer stack (talker backbone or code predictor); fields mirror the HF config attributes the
 * reference reads (talker_graph.py:36-37,63-65; predictor_graph.py:41-46). head_dim is fixed at 128. */
typedef struct {
  int32_t hidden_size;
  int32_t intermediate_size;
  int32_t num_hidden_layers;
  int32_t num_attention_heads;
  int32_t num_key_value_heads;
  int32_t vocab_size;
  float rms_norm_eps;
} fq3_stack_config;

typedef struct {
  int32_t dtype;            /* fq3_dtype */
  int32_t device;           /* CUDA ordinal */
  int32_t max_seq_len;      /* talker KV capacity (model.py:113 default 2048); <= 4096 */
  int32_t num_code_groups;  /* 16 */
  int32_t codec_eos_token_id;
  int32_t has_mtp_projection; /* small_to_mtp_projection is a Linear (1.7B) or Identity (0.6B) */
  int32_t num_ctas;         /* 0 = one CTA per SM */
  int32_t rope_positions;   /* rows of the talker cos/sin tables (>= max_seq_len + margin for rope deltas) */
  fq3_stack_config talker;
  fq3_stack_config predictor;
  int32_t max_batch;        /* request slots (KV caches + per-request state); 0/1 = one sequence, <= 32 */
} fq3_config;

/* A named tensor handed to fq3_engine_load_weights.  Names (L = layers of that stack, stacked on dim 0):
 *   t.q [L,nH*128,H]  t.k [L,nKV*128,H]  t.v  t.o [L,H,nH*128]  t.gate [L,I,H]  t.up  t.down [L,H,I]
 *   t.ln_in [L,H]  t.ln_post [L,H]  t.qnorm [L,128]  t.knorm [L,128]  t.ln_f [H]
 *   t.head [V,H] (codec_head)   t.embed [V,H] (talker.get_input_embeddings())
 *   p.* likewise for the predictor, plus
 *   p.heads [15,Vp,Hp] (lm_head[i])   p.embeds [15,Vp,Ht] (codec_embedding[i])
 *   p.mtp_w [Hp,Ht]  p.mtp_b [Hp]     (only when has_mtp_projection)
 *   t.cos t.sin [rope_positions,128]  p.cos p.sin [32,128]   -- float32 always (HF rotary tables)
 * The engine copies / repacks; the caller may free the tensors afterwards. */
typedef struct {
  const char* name;
  const void* dev_ptr;
  int64_t numel;
} fq3_tensor;

/* Sampling parameters: sampling.py:32-66 (sample_logits) + sampling.py:10-29 (apply_repetition_penalty).
 * torch.multinomial is replaced by an inverse-CDF draw on caller-supplied uniforms (noise contract). */
typedef struct {
  int32_t do_sample;
  int32_t top_k;
  float temperature;
  float top_p;
  float repetition_penalty; /* talker only; 1.0 disables */
} fq3_sampling;

/* Request state set after prefill: generate.py:120-140 (first token, past_hidden, generation_step, prefill_len,
 * rope_deltas, left-pad count from attention_mask as in talker_graph.py:172-196). */
typedef struct {
  int32_t first_token;
  int32_t prefill_len;
  int32_t gen_step;
  int32_t rope_delta;
  int32_t n_left_pad;
  int32_t max_new_tokens;
  int32_t min_new_tokens;
  int32_t trailing_len;     /* rows of trailing_text_hiddens */
} fq3_request;

/* Why the on-device loop stopped (generate.py:149-151,175-177 and the for-range bound). */
enum fq3_finish { FQ3_RUNNING = 0, FQ3_FIN_MAX_NEW = 1, FQ3_FIN_EOS = 2, FQ3_FIN_MAX_SEQ = 3 };

typedef struct {
  int32_t frames_emitted;   /* frames written by the last fq3_decode_chunk */
  int32_t finished;         /* fq3_finish */
  int32_t total_frames;     /* frames emitted since fq3_begin_request */
  int32_t next_token;       /* current cb0 token (the one the next frame would start from) */
} fq3_chunk_result;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
/* replaces TalkerGraph.__init__ / PredictorGraph.__init__ (talker_graph.py:27-59, predictor_graph.py:34-78):
 * allocates KV caches, scratch and tables.  */
int fq3_engine_create(const fq3_config* cfg, fq3_engine** out);
/* replaces holding references to the upstream nn.Modules (predictor_graph.py:53-57, talker_graph.py:41):
 * copies norm/embedding tables and repacks every GEMV weight into the per-CTA streaming tape. */
int fq3_engine_load_weights(fq3_engine* e, const fq3_tensor* tensors, int32_t n, void* stream);
void fq3_engine_destroy(fq3_engine* e);

/* Request slots.  The engine holds `max_batch` independent request slots (KV caches, predictor cache, penalty
 * bitmap, loop state).  Every per-request entry point names its slot; fq3_decode_chunk takes the list of slots
 * that advance together: one slot runs the single-sequence persistent kernel, several slots run the batched
 * kernel in which all of them share ONE pass over the weight tape per step (the reference batches left-padded
 * prompts, model.py:774-787, with per-row pad counts, talker_graph.py:177-187). */

/* ---- duck-type compatibility path (what the reference's own schedulers call) ----------------------------- */
/* TalkerGraph.prefill_kv (talker_graph.py:153-170): k,v are [n_kv, P, 128] contiguous for one layer. */
int fq3_import_kv(fq3_engine* e, int32_t slot, int32_t layer, const void* k_dev, const void* v_dev, int32_t P,
                  void* stream);
/* inverse of fq3_import_kv: cache rows [0,P) of one layer -> k,v [n_kv, P, 128] (what a caller holding the
 * reference's StaticCache would read back; used by the parity tests of the hand-written prefill) */
int fq3_export_kv(fq3_engine* e, int32_t slot, int32_t layer, void* k_dev, void* v_dev, int32_t P, void* stream);
/* TalkerGraph.set_generation_state (talker_graph.py:172-196): per-row left-pad count and rope delta. */
int fq3_set_generation_state(fq3_engine* e, int32_t slot, int32_t n_left_pad, int32_t rope_delta);
/* TalkerGraph.run (talker_graph.py:198-214): one token through 28 layers + final norm.
 * embeds_dev [H] -> hidden_out_dev [H], both model dtype. */
int fq3_talker_step(fq3_engine* e, int32_t slot, const void* embeds_dev, int32_t position, void* hidden_out_dev,
                    void* stream);
/* PredictorGraph.run (predictor_graph.py:204-214): pred_input_dev [2,H_talker] -> codes_out_dev int64[15].
 * uniforms_dev float32[15] (ignored when !do_sample). */
int fq3_predictor_run(fq3_engine* e, int32_t slot, const void* pred_input_dev, const fq3_sampling* sp,
                      const float* uniforms_dev, int64_t* codes_out_dev, void* stream);
/* sampling.py:32-66 + :10-29 on a single logits row (model dtype, [V]); history_dev int64[n_hist] or NULL.
 * suppress range is [V-1024,V) except eos when suppress_special != 0 (generate.py:46-50). token_out_dev int64[1]. */
int fq3_sample_logits(fq3_engine* e, const void* logits_dev, int32_t V, const fq3_sampling* sp, float u,
                      const int64_t* history_dev, int32_t n_hist, int32_t suppress_special, int32_t eos_id,
                      int32_t suppress_eos, int64_t* token_out_dev, void* stream);

/* ---- K3: hand-written prefill (bf16 engines) ------------------------------------------------------------------ */
/* Borrow row-major weights for the prompt GEMMs (caller keeps them alive): t.qkv [L,(nH+2nKV)*128,H] (q,k,v rows
 * concatenated), t.o [L,H,nH*128], t.gu [L,2I,H] (gate/up rows interleaved), t.down [L,H,I], t.head [V,H]. */
int fq3_engine_set_prefill_weights(fq3_engine* e, const fq3_tensor* tensors, int32_t n);
/* talker.forward prefill (generate.py:107-118) + TalkerGraph.prefill_kv (talker_graph.py:153-170) in one call:
 * embeds_dev [P,H] -> KV cache slots [0,P) of request slot `slot`, logits_out_dev [V] (codec_head on the last
 * position), hidden_out_dev [H] (post-norm hidden of the last position = past_hidden).  Positions are
 * cache index - n_left_pad (clamped at 0); keys below n_left_pad are masked. */
int fq3_prefill(fq3_engine* e, int32_t slot, const void* embeds_dev, int32_t P, int32_t n_left_pad,
                void* logits_out_dev, void* hidden_out_dev, void* stream);

/* ---- fused path (the persistent on-device loop) ---------------------------------------------------------- */
/* generate.py:120-147 / streaming.py:76-104: latch per-request state of `slot`.  past_hidden_dev [H] model dtype;
 * trailing_text_dev [trailing_len,H], tts_pad_dev [H] model dtype (borrowed until the request ends);
 * uniforms_dev float32 [(max_new_tokens+1)*16]: row s+1 = draws of frame s (col 0 talker, 1..15 predictor). */
int fq3_begin_request(fq3_engine* e, int32_t slot, const fq3_request* rq, const void* past_hidden_dev,
                      const void* trailing_text_dev, const void* tts_pad_dev, const float* uniforms_dev,
                      const fq3_sampling* sp_talker, const fq3_sampling* sp_predictor, void* stream);
/* generate.py:149-199 / streaming.py:106-173 for up to n_frames frames of every listed slot in ONE kernel launch.
 * slots[n_slots] distinct slot ids that have a latched request; codes_out_dev int64 [n_slots][n_frames][16];
 * res[n_slots] (host).  n_slots == 1: single-sequence kernel; n_slots >= 2: batched kernel, the slots advance in
 * lock-step and stop independently (EOS / max_new_tokens / max_seq_len).  Synchronises the stream. */
int fq3_decode_chunk(fq3_engine* e, const int32_t* slots, int32_t n_slots, int32_t n_frames, int64_t* codes_out_dev,
                     fq3_chunk_result* res, void* stream);
/* last post-norm talker hidden (generate.py:198 past_hidden) of `slot` -> dst_dev [H] model dtype */
int fq3_get_past_hidden(fq3_engine* e, int32_t slot, void* dst_dev, void* stream);
int fq3_max_batch(fq3_engine* e);
/* numerics probe of the batched GEMV: y[col][row] = W_seg[row,:] . x[col,:] for one weight segment of stack 0 (talker) /
 * 1 (predictor): which 0 qkv, 1 o_proj, 2 gate/up (out = model dtype [ncols][I] = silu(gate)*up), 3 down, 4 head
 * (predictor: layer = codebook).  x_dev model dtype [ncols][K]; out_dev float32 [ncols][rows] (which != 2). */
int fq3_debug_gemv(fq3_engine* e, int32_t stack, int32_t layer, int32_t which, int32_t ncols, const void* x_dev,
                   void* out_dev, void* stream);

/* ---- debugging / introspection ---------------------------------------------------------------------------- */
/* When enabled, the next talker step / predictor pass 0 dumps per-layer intermediates (float32) into an engine
 * buffer; fq3_debug_read copies `count` floats starting at `offset` to host memory.  Layout in DESIGN.md. */
int fq3_debug_enable(fq3_engine* e, int32_t on);
int fq3_debug_read(fq3_engine* e, int64_t offset, int64_t count, float* host_dst);
/* micro-benchmark: n grid barriers of flavour `kind` in one launch (tools/microbench.py) */
int fq3_barrier_test(fq3_engine* e, int32_t n, int32_t kind, void* stream);
/* bytes of packed weight tape streamed per talker step / per predictor frame (algorithmic bytes, for bench) */
int fq3_tape_bytes(fq3_engine* e, int64_t* talker_step_bytes, int64_t* predictor_frame_bytes);
int fq3_num_ctas(fq3_engine* e);
/* number of kernels launched by this engine since creation (bench.py "gpu_launches") */
int64_t fq3_launch_count(fq3_engine* e);

/* ---- K4: codec waveform decoder stack (replaces the cuDNN path under speech_tokenizer.decode, model.py:924,1093,1122)
 * geom = {device, hidden_size, decoder_dim, n_blocks, rate_0..rate_{n-1}}.  Tensor names / layouts: csrc/fq3_codec.cu.
 * fq3_codec_decode: x_dev bf16 [hidden][T4] (channels-first output of the front end) -> pcm float32 [T4*prod(rates)],
 * clamped to [-1,1]. */
typedef struct fq3_codec fq3_codec;
int fq3_codec_create(const int32_t* geom, int32_t n_geom, fq3_codec** out);
int fq3_codec_load_weights(fq3_codec* c, const fq3_tensor* tensors, int32_t n, void* stream);
int fq3_codec_decode(fq3_codec* c, const void* x_dev, int32_t T4, float* pcm_out_dev, void* stream);
/* `batch` windows of equal length in one set of launches (concurrent requests, BASELINE config 4): x_dev bf16
 * [batch][hidden][T4], pcm float32 [batch][T4*prod(rates)]; every window has its own causal left padding. */
int fq3_codec_decode_batch(fq3_codec* c, const void* x_dev, int32_t batch, int32_t T4, float* pcm_out_dev, void* stream);
/* The decoder's front end -- everything of speech_tokenizer.decode before conv_in: 16-codebook embedding mean,
 * sliding-window pre-transformer (RMSNorm, RoPE, layer scale, SwiGLU), 2 x (ConvTranspose k=s + ConvNeXt) -- as
 * hand-written kernels + the same tcgen05 GEMM.  geom = {Q, codebook_size, hidden, intermediate, n_heads, n_layers,
 * sliding_window, n_up, ratio_0 ..}; fgeom = {rms_norm_eps, rope_theta}.  Tensor names / layouts: csrc/fq3_codec.cu. */
int fq3_codec_load_frontend(fq3_codec* c, const int32_t* geom, int32_t n_geom, const float* fgeom, int32_t n_fgeom,
                            const fq3_tensor* tensors, int32_t n, void* stream);
/* speech_tokenizer.decode({"audio_codes": [batch,T,16]}) (model.py:924,1093,1122):
 * codes_dev int64 [batch][T][Q] -> pcm float32 [batch][T * total_upsample] clamped to [-1,1].  No library kernel is
 * launched.  Requires fq3_codec_load_weights + fq3_codec_load_frontend. */
int fq3_codec_decode_codes(fq3_codec* c, const int64_t* codes_dev, int32_t batch, int32_t T, float* pcm_out_dev,
                           void* stream);
/* Stateful streaming decode (replaces the reference's Phase-1 re-decode of everything so far and
 * its 25-frame Phase-2 context window, model.py:1085-1135): a stream keeps, for every causal layer, the tail of that
 * layer's input (conv history rows, the last window-1 attention keys / values), so a chunk of T frames costs T frames.
 * The PCM of a stream equals the one-shot decode of the same codes (the decoder is causal).
 * fq3_codec_stream_decode: the next T frames of n_streams distinct streams in one set of launches; codes_dev int64
 * [n_streams][T][Q]; pcm_out_dev float32 [n_streams][T * total_upsample] or NULL (state warm-up only, e.g. the ICL
 * reference frames). */
typedef struct fq3_codec_stream fq3_codec_stream;
int fq3_codec_stream_create(fq3_codec* c, fq3_codec_stream** out);
int fq3_codec_stream_reset(fq3_codec_stream* s, void* stream);
void fq3_codec_stream_destroy(fq3_codec_stream* s);
int64_t fq3_codec_stream_frames(fq3_codec_stream* s);
/* dst := src (layer histories + position), stream-ordered device copy: a stream warmed once with a voice reference is
 * the template of every later request that uses that reference */
int fq3_codec_stream_copy(fq3_codec_stream* dst, fq3_codec_stream* src, void* stream);
int fq3_codec_stream_decode(fq3_codec* c, fq3_codec_stream* const* streams, int32_t n_streams, const int64_t* codes_dev,
                            int32_t T, float* pcm_out_dev, void* stream);
double fq3_codec_flops(fq3_codec* c, int32_t T4);
double fq3_codec_frontend_flops(fq3_codec* c, int32_t T);
int64_t fq3_codec_launch_count(fq3_codec* c);
void fq3_codec_destroy(fq3_codec* c);
const char* fq3_codec_last_error(void);

/* dense-layer kernel selection for K3/K4: 0 = tcgen05 + TMA implicit GEMM, one tile per CTA, when the shape allows
 * (default); 1 = always the mma.sync kernel; 2 = persistent tcgen05 kernel with a double-buffered TMEM accumulator
 * (1, 2: A/B references). */
int fq3_set_gemm_backend(int32_t backend);

const char* fq3_last_error(void);
const char* fq3_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FQ3_ENGINE_H */
