// Internal launcher prototypes shared between the kernel translation units and the C-ABI layer.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200_whisper.h"

namespace b200w {

// one-time cudaFuncSetAttribute opt-ins (called eagerly by b200w_model_create, lazily by the launchers)
int init_gemm();
int init_attention();
int init_logmel();

// ---------------------------------------------------------------------------------------- K1 log-mel
// done_tiles == null: `out` receives log10(mel) before the clamp; else (n_audio ints) the clamp / scale is fused and `out`
// receives the normalised log-mel (cooperative launch)
int launch_logmel(const float* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                  int n_mels, const float* hann, const float* tw400, float* out, float* gmax,
                  cudaStream_t stream, int* done_tiles = nullptr);
int launch_logmel_pcm16(const int16_t* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                  int n_mels, const float* hann, const float* tw400, float* out, float* gmax,
                  cudaStream_t stream, int* done_tiles = nullptr);
int launch_logmel_finalize(float* x, const float* gmax, int n_audio, long long per_audio, cudaStream_t stream);
int launch_mel_windows(const float* mel, const float* gmax, const long long* row0, const int* size, const int* gidx,
                       int n_windows, int n_mels, __nv_bfloat16* dst, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- K5 GEMM
struct GemmParams {
  int n_batch;         // independent row slabs (M tiles never straddle slabs)
  int rows_per_batch;  // logical rows per slab; logical row = b * rows_per_batch + t
  long long out_batch_rows;  // storage rows between slabs of the output (0: rows_per_batch)
  int N;               // logical output columns (weights rows)
  int K;
  int n_store;         // columns stored (whole 32-column chunks starting below n_store are written)
  void* out;           // bf16 or f32, row-major with leading dimension ldc
  long long ldc;
  int out_f32;
  const float* bias;   // [N] or null
  const float* resid;  // f32 added after the activation, or null; row = out_row % resid_mod (0: out_row)
  long long resid_ld;
  int resid_mod;
  int gelu;
  int split_k;             // > 1: K cut into slices, slice s stores raw fp32 partials at out + s * split_stride
  long long split_stride;  // elements between partial slabs
  const char* tag;         // profile label (null: "gemm")
  // filled by the launcher
  int kb_per_split;
  int tiles_m_per_batch, tiles_n, group_m;
};

int gemm_block_n(int n_batch, int rows_per_batch, int n_store);
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int block_n, cudaStream_t stream);
// 2-CTA (cta_group::2) form for large plain GEMMs with N % 256 == 0 (gemm2.cu); B200W_GEMM2=0 disables it
int init_gemm2();
bool gemm2_enabled();
bool gemm2_applicable(const GemmParams& p);
int launch_gemm2(const void* A, long long lda, const void* W, GemmParams p, cudaStream_t stream);
int make_tmap_w(CUtensorMap* out, const void* w, int N, int K, int block_n);
int make_tmap_a(CUtensorMap* out, const void* a, int n_batch, int rows, int K, long long row_stride,
                long long batch_stride);

// ---------------------------------------------------------------------------------------- K11 decode chain
// Up to kChainMaxPhases small-M phases (GEMM with split-K slabs or bias+GELU, residual+LayerNorm) of a decoder layer
// in one cooperative launch with grid barriers between them (chain.cu).
constexpr int kChainMaxPhases = 6;
constexpr int kChainMaxGemm = 4;
constexpr int kChainGemm = 0, kChainLn = 1;
struct ChainPhase {
  int kind;
  // GEMM: out = A (rows, K) W (N, K)^T; split_k raw fp32 slabs (gelu == 0) or bias + GELU -> bf16 (gelu == 1)
  int map, tiles_n, split_k, kb_per_split, num_kb, gelu;
  void* out;
  long long ldc;
  // LayerNorm: x += bias + sum of n_split slabs of `part`; h = LN(x) * gamma + beta as bf16
  int n_split, d;
  float* x;
  const float* part;
  const float *gamma, *beta;
  __nv_bfloat16* h;
  // both
  const float* bias;
  long long split_stride;
};
struct ChainParams {
  int n_phases, n_gemm, rows;
  int tiles_m;            // 128-row tiles (1 or 2), set by chain_add_gemm
  unsigned int* counter;  // zero on entry; one per launch
  long long* timeline;    // development aid (tools/probe_chain_timeline.py): clock64 stamps, 8 per phase for 2 CTAs, or null
  ChainPhase ph[kChainMaxPhases];
};
struct ChainMaps {
  CUtensorMap a[kChainMaxGemm], b[kChainMaxGemm];
  CUtensorMap a4[kChainMaxGemm];  // A in quarter-tile boxes (cluster multicast form)
};
int init_chain();
int chain_mc_grid();  // CTAs of the cluster-multicast form that fit at once (0: form unavailable)
int chain_mc_cluster();  // CTAs per cluster of that form
int chain_add_gemm(ChainMaps* maps, ChainParams* p, const void* A, long long lda, const void* W, int N, int K, int split_k,
                   void* out, long long ldc, long long split_stride, const float* bias, bool gelu);
int chain_add_ln(ChainParams* p, float* x, const float* part, int n_split, long long split_stride, const float* bias,
                 const float* gamma, const float* beta, int d, __nv_bfloat16* h);
int launch_chain(const ChainMaps& maps, const ChainParams& p, cudaStream_t stream);
void set_chain_timeline(long long* dev, int launch_index);  // stamps go to the launch_index-th chain launch from now

// ---------------------------------------------------------------------------------------- K13 small-batch decode step
// One cooperative launch for a whole single-token decoder step of <= kSmallMaxBatch sequences (small.cu).
constexpr int kSmallMaxBatch = 5;
struct SmallArgs {
  int d, n_head, n_layer, n_vocab, B;
  const b200w_dec_layer* layers;  // DEVICE copy of the layer table
  const void* tok_emb;
  const float *dec_ln_g, *dec_ln_b;
  // decode state
  const int* pos;
  const int* finished;  // or null
  __nv_bfloat16 *k_pages, *v_pages;
  long long layer_page_stride;
  const int* block_table;
  int max_pages, page_size;
  const __nv_bfloat16* cross_kv;
  long long cross_layer_stride, cross_seq_stride;
  const int* cross_slot;
  int T;
  // workspace
  float* x;                          // (B, d) f32 residual stream, holds the embedding on entry
  __nv_bfloat16 *q, *att, *qc, *mlp;  // (B, d) x3, (B, 4d)
  float* logits;
  int logits_ld;
  float* ca_part;        // key-split cross-attention partials: (units * splits) x 66 floats
  int* ca_cnt;           // ... arrival counters per (sequence, head), zero on entry
  unsigned int* counter;  // grid-barrier counter, zero on entry
  unsigned long long* timeline;  // development aid: (2 * barriers + 2) time stamps of CTA 0, or null
};
void set_decode_small_timeline(unsigned long long* dev);
int init_decode_small();
bool decode_small_applicable(const b200w_dims& dm, int n_seq, int n_q);
int launch_decode_small(const SmallArgs& a, cudaStream_t stream);
// K13m (small_mma.cu): the same one-launch step with the projections on mma.sync, for batches of up to 16 sequences
bool decode_small_mma_applicable(const b200w_dims& dm, int n_seq, int n_q);
int launch_decode_small_mma(const SmallArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- K14 absorbed cross-attention
// Single-token decode steps: the cross-attention reads the encoder states xa (n_slots, T, d) instead of the per-layer
// K / V (absorb.cu).  q arrives as split-K slabs (+ bias) of the query projection or as bf16 (rows, d); w_ckv / b_ckv are
// the layer's fused key | value projection; `ws` is absorb_workspace_bytes() of workspace prepared once by
// absorb_prepare(); att (n_seq, d) bf16 receives what K8 would have written.
int init_absorb();
bool absorb_applicable(int n_seq, int n_head, int d, int T);
size_t absorb_workspace_bytes(int n_seq, int n_head, int d);
int absorb_prepare(void* ws, int n_seq, int n_head, int d, cudaStream_t stream);
int launch_absorbed_cross_attention(const float* q_part, int n_split, long long split_stride, const float* bias_q,
                                    const __nv_bfloat16* q_bf16, int n_seq, int n_head, const void* w_ckv,
                                    const float* b_ckv, const __nv_bfloat16* xa, int n_slots, int T, const int* slot,
                                    const int* finished, void* ws, __nv_bfloat16* att, cudaStream_t stream);
int launch_absorb_mma_bench(int m, int n, int a_mn, int ts, int reps, long long* cycles, cudaStream_t stream);
void set_absorb_timeline(long long* dev);  // development aid: 64 x 8 clock64 stamps of CTA 0, or null
int launch_absorb_probe(const __nv_bfloat16* x, const __nv_bfloat16* q, const __nv_bfloat16* p, unsigned int lbo,
                        unsigned int sbo, float* dump_s, float* dump_o, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- K4 / K10 / K9
int launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int d, __nv_bfloat16* out_bf16,
                     float* out_f32, cudaStream_t stream);
int launch_resid_ln_small(float* x, const float* part, int n_split, long long split_stride, const float* bias,
                          const float* gamma, const float* beta, int rows, int d, __nv_bfloat16* out_bf16,
                          cudaStream_t stream);
int launch_embed(const int* tokens, int tokens_ld, const int* pos, int n_seq, int n_q, const __nv_bfloat16* tok_emb,
                 const __nv_bfloat16* pos_emb, int d, int n_ctx, float* x, cudaStream_t stream);

using FilterParams = b200w_filter_params;
int launch_filter_argmax(const float* logits, const uint32_t* suppress_bits, int* tokens, int* n_tokens, int* pos,
                         float* sum_logprob, int* finished, int n_seq, const FilterParams& fp, cudaStream_t stream);
int launch_no_speech(const float* logits, int logits_ld, int n_seq, int n_vocab, int no_speech, float* out,
                     cudaStream_t stream);
int launch_language(const float* logits, int logits_ld, int n_seq, int lang_begin, int n_lang, int* lang_token,
                    float* lang_probs, cudaStream_t stream);

// ---------------------------------------------------------------------------------------- K6 / K7 / K8 attention
int launch_encoder_attention(const __nv_bfloat16* qkv, int n_batch, int T, int n_head, __nv_bfloat16* out,
                             cudaStream_t stream);
// self attention over the paged cache: appends this step's k/v rows (taken from qkv) at pos[b] + qi first.
int launch_decoder_self_attention(const __nv_bfloat16* qkv, int n_seq, int n_q, int n_head, const int* pos,
                                  __nv_bfloat16* k_pages, __nv_bfloat16* v_pages, const int* block_table,
                                  int max_pages_per_seq, int page_size, __nv_bfloat16* out, cudaStream_t stream,
                                  const float* part = nullptr, int n_split = 0, long long split_stride = 0,
                                  const float* bias = nullptr, const int* finished = nullptr);
int launch_decoder_cross_attention(const __nv_bfloat16* q, int n_seq, int n_q, int n_head,
                                   const __nv_bfloat16* cross_kv, long long seq_stride, int T, const int* slot,
                                   __nv_bfloat16* out, cudaStream_t stream, const float* part = nullptr,
                                   int n_split = 0, long long split_stride = 0, const float* bias = nullptr,
                                   const int* finished = nullptr, float* probs_out = nullptr,
                                   // small batches: workspace for the key-split form (kv_part: n_seq * n_q * n_head * 8 *
                                   // 66 floats; kv_cnt: n_seq * n_q * n_head ints, zero on first use) or null
                                   float* kv_part = nullptr, int* kv_cnt = nullptr);
int cross_attention_kv_splits(int n_seq, int n_q, int n_head);

// ---------------------------------------------------------------------------------------- K12 word alignment
// stats workspace: (2 * n_sel * n_frames + n_sel * n_tok) floats; heads: n_sel (layer-relative index, head) pairs
int launch_alignment_matrix(const float* probs, long long layer_stride, long long seq_off, int n_tok, int n_head, int T,
                            const int* heads, int n_sel, int n_frames, float* stats, float* matrix, cudaStream_t stream);
// cost: (N + 1) * (M + 1) floats, trace: as many bytes; text_idx / time_idx: N + M ints, written back to front
int launch_dtw(const float* matrix, long long ld, int N, int M, float* cost, signed char* trace, int* text_idx,
               int* time_idx, int* path_len, cudaStream_t stream);

}  // namespace b200w
