// sutro_b200 — K8: constrained decoding.
//   fsm_build_mask   for every (DFA state, token) walk the token's bytes through
//                    the byte-level DFA compiled from output_schema and pack the
//                    "token keeps the automaton alive" bits, 32 tokens per word
//                    (warp ballot).  Runs once per schema.
//   sample_greedy    per logits row: masked arg-max over the vocabulary (lowest
//                    index wins ties), append to the row's output, advance the
//                    row's DFA state by the chosen token's bytes, raise the done
//                    flag on EOS / final state / max_new_tokens.
//   prepare_decode   gathers next-step inputs (token, position, context length)
//                    from the per-slot decode state, entirely on device.
#include "common.cuh"
#include "kernels.h"

namespace sb {

namespace {

__global__ void __launch_bounds__(256)
fsm_build_mask_kernel(const int32_t* __restrict__ trans, const uint8_t* __restrict__ accept,
                      const uint8_t* __restrict__ tok_bytes, const int32_t* __restrict__ tok_off,
                      int vocab, int eos_id, uint32_t* __restrict__ mask_bits, int mask_words) {
  const int state = blockIdx.y;
  const int tok = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (tok < vocab) {
    if (tok == eos_id) {
      ok = accept[state] != 0;
    } else {
      const int b0 = tok_off[tok], b1 = tok_off[tok + 1];
      int s = state;
      for (int i = b0; i < b1 && s >= 0; ++i) s = trans[s * 256 + tok_bytes[i]];
      ok = (b1 > b0) && (s >= 0);
    }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && (tok >> 5) < mask_words)
    mask_bits[static_cast<size_t>(state) * mask_words + (tok >> 5)] = word;
}

constexpr int kSampleThreads = 512;

__global__ void __launch_bounds__(kSampleThreads) sample_greedy_kernel(SampleArgs a) {
  const int b = blockIdx.x;
  const int slot = a.row_slot[b];
  if (a.slot_done[slot]) return;
  const int state = a.slot_state ? a.slot_state[slot] : -1;
  const float* lg = a.logits + static_cast<size_t>(b) * a.ldl;
  const uint32_t* mask =
      (state >= 0 && a.mask_bits) ? a.mask_bits + static_cast<size_t>(state) * a.mask_words
                                  : nullptr;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = threadIdx.x; i < a.vocab; i += kSampleThreads) {
    if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
    const float v = lg[i];
    if (v > best || (v == best && i < best_i)) {  // NaN never wins
      best = v;
      best_i = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) {
      best = ov;
      best_i = oi;
    }
  }
  __shared__ float sv[kSampleThreads / 32];
  __shared__ int si[kSampleThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = best_i;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < kSampleThreads / 32; ++w) {
    if (sv[w] > best || (sv[w] == best && si[w] < best_i)) {
      best = sv[w];
      best_i = si[w];
    }
  }
  int tok = best_i;
  if (tok < 0 || tok >= a.vocab) tok = a.eos_id;  // nothing selectable: terminate the row

  const int row = a.slot_row[slot];
  int ngen = a.slot_ngen[slot];
  bool done = false;
  if (tok == a.eos_id && !a.ignore_eos) {
    done = true;
  } else {
    a.out_tokens[static_cast<size_t>(row) * a.out_stride + ngen] = tok;
    ++ngen;
    if (state >= 0) {
      int s = state;
      const int b0 = a.tok_off[tok], b1 = a.tok_off[tok + 1];
      for (int i = b0; i < b1 && s >= 0; ++i) s = a.fsm_trans[s * 256 + a.tok_bytes[i]];
      a.slot_state[slot] = s;
      if (s < 0 || a.fsm_final[s]) {
        done = true;
      } else if (a.fsm_tail_off != nullptr) {
        // jump-forward: everything from here to a final state is forced -> emit it now
        const int t0 = a.fsm_tail_off[s], t1 = a.fsm_tail_off[s + 1];
        if (t1 > t0) {
          const int maxnew = a.slot_maxnew[slot];
          for (int i = t0; i < t1 && ngen < maxnew; ++i)
            a.out_tokens[static_cast<size_t>(row) * a.out_stride + ngen++] = a.fsm_tail_tok[i];
          done = true;
        }
      }
    }
  }
  if (ngen >= a.slot_maxnew[slot]) done = true;
  a.slot_ngen[slot] = ngen;
  a.out_len[row] = ngen;
  a.slot_next_tok[slot] = tok;
  a.slot_pos[slot] += 1;
  if (done) a.slot_done[slot] = 1;
}

__global__ void prepare_decode_kernel(const int32_t* __restrict__ row_slot,
                                      const int32_t* __restrict__ slot_next_tok,
                                      const int32_t* __restrict__ slot_pos,
                                      int32_t* __restrict__ tok_ids, int32_t* __restrict__ tok_pos,
                                      int32_t* __restrict__ tok_slot, int32_t* __restrict__ ctx_len,
                                      int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int slot = row_slot[b];
  const int pos = slot_pos[slot];
  tok_ids[b] = slot_next_tok[slot];
  tok_pos[b] = pos;
  tok_slot[b] = slot;
  ctx_len[b] = pos + 1;
}

}  // namespace

int fsm_build_mask(const int32_t* fsm_trans, const uint8_t* fsm_accept, int n_states,
                   const uint8_t* tok_bytes, const int32_t* tok_off, int vocab, int eos_id,
                   uint32_t* mask_bits, int mask_words, cudaStream_t stream) {
  if (n_states <= 0) return 0;
  if (mask_words * 32 < vocab) {
    set_last_error("fsm_build_mask: mask_words=%d too small for vocab=%d", mask_words, vocab);
    return -1;
  }
  dim3 grid((mask_words * 32 + 255) / 256, n_states);
  fsm_build_mask_kernel<<<grid, 256, 0, stream>>>(fsm_trans, fsm_accept, tok_bytes, tok_off, vocab,
                                                  eos_id, mask_bits, mask_words);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sample_greedy(const SampleArgs& a, cudaStream_t stream) {
  if (a.B <= 0) return 0;
  sample_greedy_kernel<<<a.B, kSampleThreads, 0, stream>>>(a);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int prepare_decode(const int32_t* row_slot, const int32_t* slot_next_tok, const int32_t* slot_pos,
                   int32_t* tok_ids, int32_t* tok_pos, int32_t* tok_slot, int32_t* ctx_len, int B,
                   cudaStream_t stream) {
  if (B <= 0) return 0;
  prepare_decode_kernel<<<(B + 127) / 128, 128, 0, stream>>>(row_slot, slot_next_tok, slot_pos,
                                                             tok_ids, tok_pos, tok_slot, ctx_len, B);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace sb
