/* libb200whisper -- C ABI of the B200-native Whisper transcription hot path.
 *
 * What this replaces.  The reference repository reaches Whisper only through the console script of the
 * third-party package `mlx-whisper` (/root/reference/run:3-6); it has no FFI, operator or plugin
 * interface for this path.  The arithmetic that `mlx_whisper.transcribe` hands
 * to the MLX runtime (mlx.core ops: rfft, matmul, conv1d, layer_norm, softmax, argmax ...) is what the
 * entry points below replace, one per stage of SURVEY.md section 8a.  Each declaration cites the
 * upstream function it stands in for ("UPSTREAM" = mlx_whisper/<file>, not vendored under
 * /root/reference; restated in SURVEY.md Appendix A).
 *
 * Conventions.
 *  - Plain C types only.  Every pointer is a DEVICE pointer unless its name starts with `h_`.  The caller
 *    (PyTorch in the Python host) owns every buffer; the library allocates no persistent device memory.
 *  - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately.
 *  - Return value: 0 on success, negative b200w_status on failure; b200w_last_error() returns a
 *    thread-local description.  No global mutable state besides the launch counter and cached function
 *    attributes: safe for one host thread (or process) per GPU.
 *  - Matrices are row-major.  Weights are bf16 in the `nn.Linear` layout (out, in); biases and LayerNorm
 *    parameters are f32.  Activations feeding GEMMs are bf16, the residual stream and logits are f32.
 */
#ifndef B200_WHISPER_H_
#define B200_WHISPER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200W_ABI_VERSION 1

typedef enum {
  B200W_OK = 0,
  B200W_ERR_INVALID_ARGUMENT = -1,
  B200W_ERR_CUDA = -2,
  B200W_ERR_UNSUPPORTED = -3,
  B200W_ERR_WORKSPACE = -4,
  B200W_ERR_DRIVER = -5
} b200w_status;

const char* b200w_version(void);
const char* b200w_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports the delta). */
unsigned long long b200w_launch_count(void);
/* Per-kernel timing for benchmarks: between _begin and _end every eager launch of this library is bracketed
 * by CUDA events on its launch stream.  _end synchronises the device, writes a JSON object
 * {"<kernel>": {"launches": n, "total_ms": t}, ...} into `json` and returns the number of launches timed. */
int b200w_profile_begin(void);
int b200w_profile_end(char* h_json, size_t capacity);

/* ---------------------------------------------------------------------------------------------------
 * K1  log-mel front-end.  Replaces UPSTREAM audio.py::log_mel_spectrogram (+ stft, hanning, mel_filters);
 * reached from /root/reference/run:3.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* hann;   /* [400] periodic Hann window                                  */
  const float* tw400;  /* [25][16][2] twiddles W400^(n2*k1) as (cos, -sin)             */
} b200w_logmel_tables;
/* The mel filterbank itself (slaney scale and norm, n_mels in {80, 128}, = the reference's
 * assets/mel_filters.npz) is compiled into the kernel (csrc/mel_tables.h, tools/gen_mel_tables.py). */

/* pcm: n_audio signals of n_valid f32 samples, audio_stride samples apart.  Each signal is zero-extended
 * to n_total samples (the `padding` argument of the reference), reflect-padded by 200 and framed with
 * hop 160; out_unclamped receives (n_audio, n_total/160, n_mels) values log10(max(mel, 1e-10)) and
 * gmax[a] the maximum over signal a (the clamp reference of the reference's `log_spec.max() - 8`). */
int b200w_logmel(const float* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                 int n_mels, const b200w_logmel_tables* tables, float* out_unclamped, float* gmax, void* stream);
/* Same, from s16le PCM (the stream `load_audio` asks ffmpeg for, UPSTREAM audio.py):
 * the `/ 32768.0` of load_audio is applied while the samples are staged, so results are bit-identical to
 * b200w_logmel on the converted floats.  audio_stride in samples. */
int b200w_logmel_pcm16(const int16_t* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                       int n_mels, const b200w_logmel_tables* tables, float* out_unclamped, float* gmax, void* stream);
/* The whole of UPSTREAM log_mel_spectrogram in ONE pass over HBM: `out` receives the normalised values
 * (max(log10 mel, max_a - 8) + 4) / 4.  pcm is f32 (pcm_is_int16 == 0) or s16le.  The clamp needs the maximum over
 * the whole signal, so the kernel normalises a tile two scheduling rounds after it stored it, once every tile of the
 * signal has been counted in done_tiles (n_audio ints, scratch; zeroed by the call) -- the rows are still in L2.
 * Launched cooperatively (its CTAs wait for one another).  gmax[a] receives the maximum as in b200w_logmel. */
int b200w_logmel_normalized(const void* pcm, int pcm_is_int16, int n_audio, long long audio_stride, long long n_valid,
                            long long n_total, int n_mels, const b200w_logmel_tables* tables, float* out, float* gmax,
                            int* done_tiles, void* stream);
/* In place: x <- (max(x, gmax[a] - 8) + 4) / 4 for the per_audio values of each signal. */
int b200w_logmel_finalize(float* x, const float* gmax, int n_audio, long long per_audio, void* stream);
/* Window gather fused with the clamp/scale and the bf16 cast that feeds the conv stem.  Replaces
 * `pad_or_trim(mel[seek:seek+size], 3000).astype(fp16)` in UPSTREAM transcribe.py.  Window w takes
 * size[w] frames starting at row row0[w] of `mel` (rows of n_mels f32), clamps with gmax[gidx[w]] and
 * writes rows 1..3000 of its (3002, n_mels) bf16 slab in dst; rows 0 and 3001 are zero (conv padding).
 * gmax == NULL: `mel` already holds normalised log-mel values; only the cast and the padding are applied. */
int b200w_mel_windows(const float* mel, const float* gmax, const long long* row0, const int* size, const int* gidx,
                      int n_windows, int n_mels, void* dst_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K5  bf16 tensor-core GEMM, C = epilogue(A W^T).  Replaces mlx.nn.Linear / as_linear in UPSTREAM
 * whisper.py (MultiHeadAttention.{query,key,value,out}, mlp1, mlp2, token_embedding.as_linear).
 * ------------------------------------------------------------------------------------------------- */
#define B200W_GEMM_GELU 1    /* exact (erf) GELU after the bias                    */
#define B200W_GEMM_OUT_F32 2 /* C is f32 (otherwise bf16)                          */
/* A: (M, K) bf16 with row stride lda elements; W: (N, K) bf16; bias: [N] f32 or NULL; resid: (M, ldc) f32
 * added after the activation or NULL (may alias C); C: (M, ldc).  Requires K % 8 == 0, ldc >= N rounded up
 * to 32; columns up to that rounding are written. */
int b200w_gemm_bf16(const void* A, long long lda, const void* W, void* C, long long ldc, const float* bias,
                    const float* resid, int M, int N, int K, int flags, void* stream);

/* K14: single-token cross-attention in absorbed form -- reads the encoder states xa ((n_slots, T, d) bf16, d = 64 *
 * n_head) instead of the cached per-layer K / V: scores = (Wk_h^T q_h) . xa_t, output = Wv_h (sum_t p_t xa_t) + bv_h,
 * both contractions on tcgen05.  q: (n_seq, d) bf16 query projection; w_ckv / b_ckv: the layer's fused key | value
 * projection ((2d, d) bf16, [2d] f32); slot: sequence -> row of xa; finished: sequences to skip (may be NULL);
 * out: (n_seq, d) bf16.  Same result as b200w_decoder_cross_attention on K / V = xa W^T + b up to 16-bit rounding.
 * Replaces UPSTREAM whisper.py::MultiHeadAttention.__call__ (cross-attention branch with kv_cache) for n_q == 1. */
size_t b200w_absorbed_cross_attention_workspace_bytes(int n_seq, int n_head);
int b200w_absorbed_cross_attention(const void* q, int n_seq, int n_head, const void* w_ckv, const float* b_ckv,
                                   const void* xa, int n_slots, int T, const int* slot, const int* finished,
                                   void* workspace, size_t workspace_bytes, void* out, void* stream);

/* Split-K form for decode steps (M <= 128 rows, one row tile): the K range is cut into
 * b200w_gemm_splitk_slices(K, split_k) slices so that all SMs stream weights; slice s stores its raw fp32
 * partial product to part + s * split_stride (rows of ldp floats).  No bias / activation: the consumer
 * (b200w_residual_layernorm, b200w_decoder_*_attention_splitk) sums the slabs. */
int b200w_gemm_bf16_splitk(const void* A, long long lda, const void* W, float* part, long long ldp, long long split_stride,
                           int M, int N, int K, int split_k, void* stream);
int b200w_gemm_splitk_slices(int K, int split_k);

/* K2/K3  conv stem as implicit GEMM.  Replaces nn.Conv1d(k=3, padding=1, stride) + nn.gelu (+ positional
 * add) in UPSTREAM whisper.py::AudioEncoder.__call__.  x_padded: (B, T_in + 2, C_in) bf16 with zero first
 * and last rows; w: (C_out, 3 * C_in) bf16 = the MLX (out, k, in) layout flattened; out: (B * T_out, C_out)
 * with T_out = T_in / stride, bf16 or f32 (out_f32); pos: optional (T_out, C_out) f32 added after the GELU. */
int b200w_conv1d_gelu(const void* x_padded, const void* w, const float* bias, int n_batch, int t_in, int c_in,
                      int c_out, int stride, const float* pos, void* out, long long out_ld, int out_f32, void* stream);

/* K4  LayerNorm (eps 1e-5, affine) over rows of f32 x; writes bf16 and/or f32 (either may be NULL).
 * Replaces nn.LayerNorm in UPSTREAM whisper.py::ResidualAttentionBlock. */
int b200w_layernorm(const float* x, const float* gamma, const float* beta, int rows, int d, void* out_bf16,
                    float* out_f32, void* stream);

/* K4b decode-step residual update fused with the next LayerNorm (rows <= a few hundred, d % 128 == 0):
 *   x <- x + bias + sum_{s < n_split} part[s]   (n_split == 0: x unchanged),   out <- LayerNorm(x) as bf16.
 * Replaces `x = x + out_proj(...)` / `x = x + mlp2(...)` followed by the next nn.LayerNorm in UPSTREAM
 * whisper.py::ResidualAttentionBlock for single-token decoder steps. */
int b200w_residual_layernorm(float* x, const float* part, int n_split, long long split_stride, const float* bias,
                             const float* gamma, const float* beta, int rows, int d, void* out_bf16, void* stream);

/* K6  encoder self-attention (non-causal) on the fused (B*T, 3d) bf16 QKV activation -> (B*T, d) bf16.
 * Replaces UPSTREAM whisper.py::MultiHeadAttention.qkv_attention for the AudioEncoder. */
int b200w_encoder_attention(const void* qkv, int n_batch, int T, int n_head, void* out, void* stream);

/* K7  decoder self-attention over the paged KV cache.  qkv: (n_seq * n_q, 3d) bf16 for the n_q new tokens
 * of each sequence; their k/v rows are appended at positions pos[b] .. pos[b]+n_q-1 of the pages named by
 * block_table (n_seq, max_pages); query qi attends causally to positions <= pos[b] + qi.  Replaces the
 * `kv_cache` concatenate + masked qkv_attention of UPSTREAM whisper.py::TextDecoder. */
int b200w_decoder_self_attention(const void* qkv, int n_seq, int n_q, int n_head, const int* pos, void* k_pages,
                                 void* v_pages, const int* block_table, int max_pages, int page_size, void* out,
                                 void* stream);

/* K7 with the fused QKV projection given as split-K partial slabs (rows of 3d floats, one new token per
 * sequence): reduces q/k/v (+ bias_qkv), rounds to bf16, appends k/v to the pages, attends. */
int b200w_decoder_self_attention_splitk(const float* qkv_part, int n_split, long long split_stride, const float* bias_qkv,
                                        int n_seq, int n_head, const int* pos, void* k_pages, void* v_pages,
                                        const int* block_table, int max_pages, int page_size, void* out, void* stream);

/* K8  decoder cross-attention.  q: (n_seq * n_q, d) bf16; cross_kv: slots of (T, 2d) bf16 rows [K | V],
 * seq_stride elements apart; slot[b] names the slot of sequence b.  Replaces the cached-`xa` branch of
 * UPSTREAM whisper.py::MultiHeadAttention.__call__. */
int b200w_decoder_cross_attention(const void* q, int n_seq, int n_q, int n_head, const void* cross_kv,
                                  long long seq_stride, int T, const int* slot, void* out, void* stream);

/* K8 with the query projection given as split-K partial slabs (rows of d floats, one query per sequence). */
int b200w_decoder_cross_attention_splitk(const float* q_part, int n_split, long long split_stride, const float* bias_q,
                                         int n_seq, int n_head, const void* cross_kv, long long seq_stride, int T,
                                         const int* slot, void* out, void* stream);

/* K10 token + positional embedding: x[b*n_q+qi] = tok_emb[tokens[b][pos[b]+qi]] + pos_emb[pos[b]+qi] (f32). */
int b200w_embed(const int* tokens, int tokens_ld, const int* pos, int n_seq, int n_q, const void* tok_emb,
                const void* pos_emb, int d, int n_ctx, float* x, void* stream);

/* K9  fused logit filter + log-softmax + token selection.  Replaces UPSTREAM decoding.py::SuppressBlank,
 * SuppressTokens, ApplyTimestampRules and GreedyDecoder.update (which run on the host with a sync per
 * step).  For each sequence b with history tokens[b][0..n_tokens[b]): picks the next token from
 * logits[b] (greedy, or Gumbel-max sampling when temperature > 0), appends it, adds its log-probability
 * to sum_logprob[b] unless the sequence already ended, forces EOT after EOT, sets finished[b],
 * pos[b] = old n_tokens[b] and n_tokens[b] += 1. */
typedef struct {
  int n_vocab;
  int logits_ld;
  int sample_begin;
  int eot, blank, no_timestamps, timestamp_begin, no_speech;
  int max_initial_timestamp_index; /* < 0 disables the option */
  int apply_timestamp_rules;       /* 0 for without_timestamps */
  int suppress_blank;
  int tokens_ld;
  float temperature;
  unsigned long long seed;
} b200w_filter_params;
int b200w_filter_argmax(const float* logits, const uint32_t* suppress_bits, int* tokens, int* n_tokens, int* pos,
                        float* sum_logprob, int* finished, int n_seq, const b200w_filter_params* fp, void* stream);
/* softmax(logits[b])[no_speech] -> out[b]   (UPSTREAM decoding.py::DecodingTask._main_loop, step 0) */
int b200w_no_speech_prob(const float* logits, int logits_ld, int n_seq, int n_vocab, int no_speech, float* out,
                         void* stream);
/* argmax / softmax restricted to the language tokens (UPSTREAM decoding.py::detect_language) */
int b200w_detect_language(const float* logits, int logits_ld, int n_seq, int lang_begin, int n_lang, int* lang_token,
                          float* lang_probs, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Engine: the layer loops of UPSTREAM whisper.py::AudioEncoder / TextDecoder sequenced in native code.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
  int n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer;
} b200w_dims;

typedef struct {
  const float *attn_ln_g, *attn_ln_b;
  const void* w_qkv;   /* (3d, d): query | key | value */
  const float* b_qkv;  /* [3d], key slice zero */
  const void* w_out;
  const float* b_out;
  const float *mlp_ln_g, *mlp_ln_b;
  const void* w_mlp1;
  const float* b_mlp1;
  const void* w_mlp2;
  const float* b_mlp2;
} b200w_enc_layer;

typedef struct {
  const float *attn_ln_g, *attn_ln_b;
  const void* w_qkv;
  const float* b_qkv;
  const void* w_out;
  const float* b_out;
  const float *cross_ln_g, *cross_ln_b;
  const void* w_cq;    /* cross query (d, d) */
  const float* b_cq;
  const void* w_ckv;   /* cross key | value (2d, d), applied to the encoder states */
  const float* b_ckv;  /* [2d], key slice zero */
  const void* w_cout;
  const float* b_cout;
  const float *mlp_ln_g, *mlp_ln_b;
  const void* w_mlp1;
  const float* b_mlp1;
  const void* w_mlp2;
  const float* b_mlp2;
} b200w_dec_layer;

typedef struct {
  b200w_dims dims;
  const void* conv1_w;   /* (d, 3*n_mels) */
  const float* conv1_b;
  const void* conv2_w;   /* (d, 3*d) */
  const float* conv2_b;
  const float* enc_pos;  /* (n_audio_ctx, d) f32 sinusoids */
  const float *ln_post_g, *ln_post_b;
  const b200w_enc_layer* h_enc_layers; /* HOST array [n_audio_layer] */
  const void* tok_emb;   /* (n_vocab, d) bf16, also the tied output projection */
  const void* dec_pos;   /* (n_text_ctx, d) bf16 */
  const float *dec_ln_g, *dec_ln_b;
  const b200w_dec_layer* h_dec_layers; /* HOST array [n_text_layer] */
} b200w_weights;

typedef struct b200w_model b200w_model;
/* Copies the pointer tables (not the weights). */
int b200w_model_create(const b200w_weights* w, b200w_model** out);
void b200w_model_destroy(b200w_model* m);

size_t b200w_encoder_workspace_bytes(const b200w_model* m, int n_windows);
/* mel_padded: (n_windows, 3002, n_mels) bf16 from b200w_mel_windows.  Writes the encoder states
 * ln_post(x) as bf16 (n_windows * n_audio_ctx, d) (the operand of the cross K/V projections) and, if
 * xa_f32 != NULL, as f32.  Replaces UPSTREAM whisper.py::AudioEncoder.__call__.
 * stop_after_layers < 0 runs the full stack; k >= 0 stops after k blocks and stores the f32 residual
 * stream to xa_f32 without ln_post (parity probes). */
int b200w_encoder_forward(const b200w_model* m, const void* mel_padded, int n_windows, void* workspace,
                          size_t workspace_bytes, void* xa_bf16, float* xa_f32, int stop_after_layers, void* stream);

/* Cross K/V projections of all decoder layers for n_windows encoder states:
 * cross_kv[layer][slot0 + w] = [xa_w W_k^T | xa_w W_v^T + b_v], layers layer_stride elements apart, slots
 * n_audio_ctx * 2d elements apart.  Replaces the first-call branch of cross attention in UPSTREAM
 * whisper.py::MultiHeadAttention.__call__ (`xa` given, cache empty). */
int b200w_cross_kv(const b200w_model* m, const void* xa_bf16, int n_windows, void* cross_kv, long long layer_stride,
                   int slot0, void* stream);

typedef struct {
  int n_seq;
  int* tokens;          /* (n_seq, tokens_ld) */
  int tokens_ld;
  int* n_tokens;        /* (n_seq) tokens present          */
  int* pos;             /* (n_seq) tokens already cached   */
  float* sum_logprob;   /* (n_seq) */
  int* finished;        /* (n_seq) */
  float* no_speech;     /* (n_seq) written when sot_index >= 0 */
  void* k_pages;        /* (n_text_layer, n_pages, page_size, d) bf16 */
  void* v_pages;
  long long layer_page_stride; /* elements between layers */
  int* block_table;     /* (n_seq, max_pages) */
  int max_pages, page_size;
  const void* cross_kv; /* (n_text_layer, n_slots, n_audio_ctx, 2d) bf16 */
  long long cross_layer_stride;
  int* cross_slot;      /* (n_seq) */
  float* logits;        /* (n_seq, logits_ld) f32: logits of the last new token */
  float* logits_aux;    /* (n_seq, logits_ld) f32: logits at sot_index (prompt step) */
  int logits_ld;        /* >= n_vocab rounded up to 128 */
  const uint32_t* suppress_bits; /* ceil(n_vocab / 32) words, bit v set = token v suppressed */
  const void* xa;       /* (xa_slots, n_audio_ctx, d) bf16 encoder states behind cross_kv, or NULL: when given, single-token
                         * steps of batches >= 16 run the cross-attention in absorbed form (K14) and stream xa instead
                         * of cross_kv (B200W_ABSORB=0 keeps K8) */
  int xa_slots;
} b200w_decode_state;

size_t b200w_decoder_workspace_bytes(const b200w_model* m, int n_seq, int n_q);
/* One decoder forward over the n_q newest tokens of every sequence followed by token selection:
 * embed -> n_text_layer x (LN, QKV, paged self-attention, out+residual, LN, Q, cross-attention,
 * out+residual, LN, MLP) -> LN -> tied logits GEMM (last token only) -> K9.  sot_index >= 0 (prompt
 * step) additionally evaluates the logits at that prompt position for no_speech.  With select == 0 the
 * K9 stage is skipped (teacher forcing / language id: the caller reads `logits`).
 * Replaces UPSTREAM decoding.py::Inference.logits + the filter / GreedyDecoder.update half of _main_loop. */
int b200w_decoder_step(const b200w_model* m, const b200w_decode_state* st, int n_q, int sot_index, int select,
                       const b200w_filter_params* fp, void* workspace, size_t workspace_bytes, void* stream);

/* -------------------------------------------------------------------------------------------------
 * Word-level timestamps (UPSTREAM mlx_whisper/timing.py)
 * ------------------------------------------------------------------------------------------------- */
/* Teacher-forced forward over the n_q loaded tokens of every sequence (the decoder half of UPSTREAM
 * whisper.py::Whisper.forward_with_cross_qk): the logits of EVERY position go to logits_all
 * ((n_seq * n_q, logits_ld) f32, may be NULL) and the cross-attention probabilities of the layers
 * >= probs_first_layer to cross_probs ((n_text_layer - probs_first_layer, n_seq, n_q, n_text_head, 1500) f32,
 * may be NULL).  Same workspace as b200w_decoder_step(n_seq, n_q). */
int b200w_decoder_forward_full(const b200w_model* m, const b200w_decode_state* st, int n_q, void* workspace,
                               size_t workspace_bytes, float* logits_all, float* cross_probs, int probs_first_layer,
                               void* stream);
/* timing.py::find_alignment's matrix for sequence `seq`: the probabilities of the n_sel selected (layer - first
 * layer, head) pairs `heads` (device, 2 * n_sel ints) are renormalised over the first n_frames frames, normalised
 * over the token axis ((w - mean) / std), median-filtered (width 7, reflect padding) along frames and averaged over
 * the heads -> matrix (n_q, n_frames) f32.  stats: workspace of 2 * n_sel * n_frames + n_sel * n_q floats. */
int b200w_alignment_matrix(const float* cross_probs, int n_layers_stored, int n_seq, int seq, int n_q, int n_head,
                           int n_ctx, const int* heads, int n_sel, int n_frames, float* stats, float* matrix,
                           void* stream);
/* timing.py::dtw on -matrix (N rows of M columns, row stride ld): monotonic alignment path, written BACK TO FRONT to
 * text_idx / time_idx (N + M ints each) with its length in path_len (device).  cost: (N + 1) * (M + 1) floats and
 * trace: as many bytes of workspace. */
int b200w_dtw(const float* matrix, long long ld, int N, int M, float* cost, signed char* trace, int* text_idx,
              int* time_idx, int* path_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_WHISPER_H_ */
