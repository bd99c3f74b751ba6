/* sutro_b200 — C-ABI of the B200-native local backend for sutro.infer().
 *
 * The reference SDK (sutro-sh/sutro) has no FFI: Sutro.infer() serialises the
 * column and POSTs it to a hosted service (sutro/sdk.py:195-223).  This header
 * is therefore the boundary a maintainer would bind *instead of* that POST; each
 * entry point names the reference interface it stands in for.  INTEGRATION.md
 * shows the ctypes stub that replaces `do_request("POST", "batch-inference")`.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from sb200_last_error() (thread-local, never NULL);
 *   - nothing throws across the boundary;
 *   - "device pointer" arguments are raw CUDA device addresses (PyTorch tensors'
 *     data_ptr() on the Python side); `stream` is a cudaStream_t passed as void*;
 *   - strings travel Arrow-style: one byte blob + int64 offsets[n+1];
 *   - the caller owns inputs; the engine owns outputs until sb200_result_free().
 */
#ifndef SUTRO_B200_H_
#define SUTRO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 1

const char* sb200_last_error(void);
int sb200_abi_version(void);
int sb200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ------------------------------------------------------------------------
 * Kernel-level entry points (used by the parity tests; the engine calls the
 * same launchers internally).  Shapes are in elements, tensors are row-major.
 * ---------------------------------------------------------------------- */

/* epilogue selectors for sb200_gemm_bf16_tn */
#define SB200_EPI_STORE_BF16 0    /* D = bf16(A W^T)                              */
#define SB200_EPI_RESIDUAL_BF16 1 /* D = bf16(bf16(A W^T) + R)                    */
#define SB200_EPI_SWIGLU_BF16 2   /* D[:,j] = silu(acc[:,2j]) * acc[:,2j+1]       */
#define SB200_EPI_STORE_F32 3     /* D = A W^T in fp32 (logits)                   */

/* K1: D[M,N] = A[M,K] W[N,K]^T on tcgen05 tensor cores.  A may be the head of a
 * larger buffer with a_rows >= M rows.  block_n: 0 = auto, else 64/128/256. */
int sb200_gemm_bf16_tn(const void* a, int a_rows, const void* w, void* d, const void* resid, int M,
                       int N, int K, int ldd, int epilogue, int block_n, void* stream);

/* K4: out = w * bf16(x * rsqrt(mean(x^2)+eps))  (transformers Qwen3RMSNorm) */
int sb200_rmsnorm(const void* x, const void* w, void* out, int rows, int d, float eps,
                  void* stream);

/* K9: embedding gather and the embedding-model head (L2-normalised fp32 rows) */
int sb200_embed_gather(const int32_t* ids, const void* table, void* out, int rows, int d,
                       void* stream);
int sb200_l2_normalize_rows(const void* x, float* out, int rows, int d, void* stream);

/* K5: per-head q/k RMSNorm (NULL weights = none) + RoPE, K/V scatter into the
 * paged cache pool[page][kv_head][K|V][16][128] (swizzled). */
int sb200_rope_kv_write(void* qkv, const void* q_norm_w, const void* k_norm_w, const void* cos_tab,
                        const void* sin_tab, const int32_t* tok_slot, const int32_t* tok_pos,
                        const int32_t* page_table, int max_pages, void* kv_layer, int T, int hq,
                        int hkv, float eps, void* stream);

/* K2: paged-KV decode attention, one query token per sequence. */
int sb200_attn_decode(const void* qkv, void* out, const void* kv_layer, const int32_t* page_table,
                      int max_pages, const int32_t* row_slot, const int32_t* ctx_len, int B, int hq,
                      int hkv, float scale, void* stream);

/* K3: causal varlen prefill attention over the paged cache. */
int sb200_attn_prefill(const void* qkv, void* out, const void* kv_layer, const int32_t* page_table,
                       int max_pages, const int32_t* work, int n_work, const int32_t* seq_slot,
                       const int32_t* seq_q_start, const int32_t* seq_q_len,
                       const int32_t* seq_past, int hq, int hkv, float scale, void* stream);
int sb200_attn_prefill_q_tile(int hq, int hkv);

/* K8: token-level mask of a byte DFA compiled from output_schema. */
int sb200_fsm_build_mask(const int32_t* fsm_trans, const uint8_t* fsm_accept, int n_states,
                         const uint8_t* tok_bytes, const int32_t* tok_off, int vocab, int eos_id,
                         uint32_t* mask_bits, int mask_words, void* stream);

/* K7: batched byte-level BPE on the GPU.  merges:[n,2] left/right ids in rank order
 * (merged_ids NULL => id 256+rank); cls_table: 0x110000 bytes, Unicode class per code
 * point (0 other, 1 letter, 2 number, 3 white space); digits: max digits per pre-token
 * (1 Qwen, 3 Llama-3); tok_bytes/tok_off: vocabulary byte strings (host pointers).
 * encode/decode take DEVICE pointers; out_tokens_dev needs capacity n_bytes;
 * decode with out_bytes_dev == NULL only fills row_byte_off_dev (sizing pass). */
int sb200_tokenizer_create(const int32_t* merges, int n_merges, const int32_t* merged_ids,
                           const uint8_t* cls_table, int digits, const uint8_t* tok_bytes,
                           const int32_t* tok_off, int vocab, void** out);
void sb200_tokenizer_destroy(void* tok);
int sb200_tokenizer_encode(void* tok, const uint8_t* text_dev, int64_t n_bytes,
                           const int64_t* row_off_dev, int64_t n_rows, int32_t* out_tokens_dev,
                           int64_t* row_tok_off_dev, void* stream);
int sb200_tokenizer_decode(void* tok, const int32_t* toks_dev, int64_t n_tok,
                           const int64_t* row_tok_off_dev, int64_t n_rows, uint8_t* out_bytes_dev,
                           int64_t* row_byte_off_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUTRO_B200_H_ */
