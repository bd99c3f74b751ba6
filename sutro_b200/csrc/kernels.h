// sutro_b200 — internal launcher declarations shared by the kernels, the engine
// and the C-ABI (capi.cu).  All launchers return 0 on success, -1 on error
// (message retrievable through sb::last_error()).  No launcher synchronises.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

const char* last_error();
void set_last_error(const char* fmt, ...);

// ---------------------------------------------------------------------------
// geometry of the paged KV cache (see DESIGN.md "data layout in HBM")
//   pool[layer][page][kv_head][K|V][kPageTokens][kHeadDim]  bf16
//   inside one [16][128] tile the 16-byte chunk c of token row r is stored at
//   chunk position c ^ (r & 7): a flat 4 KiB bulk copy then lands in shared
//   memory already bank-conflict-free for ldmatrix.
// ---------------------------------------------------------------------------
constexpr int kHeadDim = 128;
constexpr int kPageTokens = 16;
constexpr int kTileElems = kPageTokens * kHeadDim;   // one K (or V) tile of one head
constexpr int kTileBytes = kTileElems * 2;           // 4096

enum GemmEpilogue {
  EPI_STORE_BF16 = 0,
  EPI_RESIDUAL_BF16 = 1,
  EPI_SWIGLU_BF16 = 2,
  EPI_STORE_F32 = 3,
  EPI_QKV_ROPE = 4,  // QKV projection fused with q/k-norm + RoPE + paged K/V write (K1+K5)
};

// Extra operands of EPI_QKV_ROPE: everything rope_kv_write needs, applied to the
// accumulators in the GEMM epilogue so the [T, qkv_dim] tensor never round-trips HBM.
struct QkvEpiArgs {
  const int32_t* tok_pos;
  const int32_t* tok_slot;
  const int32_t* page_table;
  int max_pages;
  void* kv_layer;
  const void* cos_tab;
  const void* sin_tab;
  const void* q_norm_w;  // nullptr: no per-head norm
  const void* k_norm_w;
  int hq, hkv;
  float eps;
  int write_dense;  // also leave post-norm/RoPE K and V in their qkv columns (dense prefill attention)
};

// K1 — D = A · W^T (+ fused epilogue); A:[a_rows>=M, K] bf16, W:[N,K] bf16.
int gemm_bf16_tn(const void* a, int a_rows, const void* w, void* d, const void* resid, int M,
                 int N, int K, int ldd, int epilogue, int block_n, cudaStream_t stream,
                 const QkvEpiArgs* qkv_args = nullptr);
int gemm_pick_block_n(int M, int N);

// K4 — RMSNorm over rows: out = bf16(w * bf16(x * rsqrt(mean(x^2) + eps))).
int rmsnorm(const void* x, const void* w, void* out, int rows, int d, float eps,
            cudaStream_t stream);

// K9a — embedding gather: out[t,:] = table[ids[t],:].
int embed_gather(const int32_t* ids, const void* table, void* out, int rows, int d,
                 cudaStream_t stream);
// gather rows of x by index (last-token selection before the lm_head).
int gather_rows(const int32_t* idx, const void* x, void* out, int rows, int d,
                cudaStream_t stream);

// K5 — per-head q/k RMSNorm (optional) + RoPE on q and k, K/V scatter into the
// paged cache.  qkv:[T, (hq+2*hkv)*128] is updated in place for q.
int rope_kv_write(void* qkv, const void* q_norm_w, const void* k_norm_w, const void* cos_tab,
                  const void* sin_tab, const int32_t* tok_slot, const int32_t* tok_pos,
                  const int32_t* page_table, int max_pages, void* kv_layer, int T, int hq,
                  int hkv, float eps, cudaStream_t stream);

// K2 — paged-KV decode attention (one query token per sequence).
int attn_decode(const void* qkv, void* out, const void* kv_layer, const int32_t* page_table,
                int max_pages, const int32_t* row_slot, const int32_t* ctx_len, int B, int hq,
                int hkv, float scale, cudaStream_t stream);

void attn_decode_force_variant(int v);  // 0 auto, 1 split kernel, 2 warp-per-pair kernel

// K3 — causal prefill attention over the paged cache (varlen batch).
// work: [n_work] {seq, q_tile_start}; per-seq arrays are indexed by seq.
int attn_prefill(const void* qkv, void* out, const void* kv_layer, const int32_t* page_table,
                 int max_pages, const int32_t* work, int n_work, const int32_t* seq_slot,
                 const int32_t* seq_q_start, const int32_t* seq_q_len, const int32_t* seq_past,
                 int hq, int hkv, float scale, cudaStream_t stream);
int attn_prefill_q_tile(int hq, int hkv);

// K3 (tcgen05) — causal prefill attention over DENSE K/V: new tokens' K/V are the k/v columns
// of the qkv buffer itself (post-norm/RoPE, written by the fused QKV epilogue), the shared
// prefix's K/V live in prefix_kv[n_layers][prefix_rows][2*hkv*128] (K heads | V heads).
// items: [n_items] {seq, q_tile_start} with q tile = attn_prefill_q_tile(); every sequence
// attends to prefix rows [0, seq_past) and its own rows causally.
int attn_prefill_dense(const void* qkv, int t_rows, void* out, const void* prefix_kv,
                       int prefix_rows, int n_layers, int layer, const int32_t* items, int n_items,
                       const int32_t* seq_q_start, const int32_t* seq_q_len,
                       const int32_t* seq_past, int hq, int hkv, float scale, cudaStream_t stream);

// bf16 tensor map, 128-byte swizzle, rank 2..5 (gemm_tcgen05.cu)
int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

// K8 — FSM token-mask build, masked greedy sampling + FSM advance.
struct SampleArgs {
  const float* logits;      // [B, ldl]
  int ldl;
  int vocab;
  int B;
  const int32_t* row_slot;  // [B] slot of each logits row
  // per-slot decode state
  int32_t* slot_state;      // FSM state (or -1 when unconstrained)
  int32_t* slot_ngen;       // tokens generated so far
  int32_t* slot_next_tok;   // out: sampled token (input of next step)
  int32_t* slot_pos;        // position of next token (incremented)
  int32_t* slot_done;       // out: 1 when the row finished at this step
  const int32_t* slot_row;  // global row id of the slot
  const int32_t* slot_maxnew;
  int32_t* out_tokens;      // [n_rows, out_stride]
  int32_t* out_len;         // [n_rows]
  int out_stride;
  // FSM tables (null when no schema)
  const uint32_t* mask_bits;   // [n_states, mask_words]
  int mask_words;
  const int32_t* fsm_trans;    // [n_states, 256] byte transitions, -1 = dead
  const uint8_t* fsm_accept;   // [n_states]
  const uint8_t* fsm_final;    // [n_states] accepting with no outgoing edge
  const int32_t* fsm_tail_off; // [n_states+1] or null: forced terminal continuation per state
  const int32_t* fsm_tail_tok;
  const uint8_t* tok_bytes;    // vocab byte blob
  const int32_t* tok_off;      // [vocab+1]
  int eos_id;
  int ignore_eos;
  // sampling_params (reference: opaque dict forwarded to the service, sutro/sdk.py:203)
  float temperature;       // 0 = greedy (arg-max, lowest index wins ties)
  int top_k;               // <= 0: off
  float top_p;             // >= 1: off
  uint64_t seed;
  int seed_per_row;        // random_seed_per_input (sutro/sdk.py:204): 1 = own stream per row
  const int64_t* row_ids;  // optional [n_rows]: stream id per row (null = row index)
  float* slot_cum_logprob; // nullable: running sum of log p(token) under the masked softmax
  float* out_cum_logprob;  // [n_rows], written together with out_len
};
int sample_greedy(const SampleArgs& a, cudaStream_t stream);  // dispatches on a.temperature
int prepare_decode(const int32_t* row_slot, const int32_t* slot_next_tok, const int32_t* slot_pos,
                   int32_t* tok_ids, int32_t* tok_pos, int32_t* tok_slot, int32_t* ctx_len, int B,
                   cudaStream_t stream);
int fsm_build_mask(const int32_t* fsm_trans, const uint8_t* fsm_accept, int n_states,
                   const uint8_t* tok_bytes, const int32_t* tok_off, int vocab, int eos_id,
                   uint32_t* mask_bits, int mask_words, cudaStream_t stream);

// K7 — GPU byte-level BPE tokenizer / detokenizer (tokenizer.cu).
struct Tokenizer;
int tokenizer_create(const int32_t* merges, int n_merges, const int32_t* merged_ids,
                     const uint8_t* cls_table, int digits, const uint8_t* tok_bytes,
                     const int32_t* tok_off, int vocab, Tokenizer** out);
int tokenizer_set_word_overrides(Tokenizer* t, const int32_t* seq_tokens, const int32_t* seq_off,
                                 const int32_t* ids, int n);
void tokenizer_destroy(Tokenizer* t);
const uint8_t* tokenizer_tok_bytes(const Tokenizer* t);
const int32_t* tokenizer_tok_off(const Tokenizer* t);
int tokenizer_encode(Tokenizer* t, const uint8_t* text_dev, int64_t n_bytes,
                     const int64_t* row_off_dev, int64_t n_rows, int32_t* out_tokens_dev,
                     int64_t* row_tok_off_dev, cudaStream_t stream);
int tokenizer_decode(Tokenizer* t, const int32_t* toks_dev, int64_t n_tok,
                     const int64_t* row_tok_off_dev, int64_t n_rows, uint8_t* out_bytes_dev,
                     int64_t* row_byte_off_dev, cudaStream_t stream);

// K9b — embedding head: last-token rows -> fp32 L2-normalised vectors.
int l2_normalize_rows(const void* x, float* out, int rows, int d, cudaStream_t stream);

}  // namespace sb
