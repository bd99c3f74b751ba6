// sutro_b200 — K8: constrained decoding.
//   fsm_build_mask   for every (DFA state, token) walk the token's bytes through
//                    the byte-level DFA compiled from output_schema and pack the
//                    "token keeps the automaton alive" bits, 32 tokens per word
//                    (warp ballot).  Runs once per schema.
//   sample_greedy    per logits row: masked arg-max over the vocabulary (lowest
//                    index wins ties), append to the row's output, advance the
//                    row's DFA state by the chosen token's bytes, raise the done
//                    flag on EOS / final state / max_new_tokens.
//   sample_random    temperature / top-k / top-p sampling under the same mask
//                    (exact radix-select thresholds, Philox4x32-10 draws).
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
constexpr int kSampleWarps = kSampleThreads / 32;

// Row bookkeeping shared by the greedy and the stochastic sampler (thread 0 of the CTA):
// append the token, advance the DFA by its bytes, apply jump-forward tails, raise `done`.
__device__ void finish_row(const SampleArgs& a, int slot, int state, int tok, float logprob) {
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
  if (a.slot_cum_logprob != nullptr) {
    const float c = a.slot_cum_logprob[slot] + logprob;
    a.slot_cum_logprob[slot] = c;
    if (a.out_cum_logprob != nullptr) a.out_cum_logprob[row] = c;
  }
  a.slot_next_tok[slot] = tok;
  a.slot_pos[slot] += 1;
  if (done) a.slot_done[slot] = 1;
}

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax x, ArgMax y) {
  return (y.v > x.v || (y.v == x.v && y.i < x.i)) ? y : x;  // NaN never wins
}
__device__ ArgMax block_argmax(ArgMax m, float* sv, int* si) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    m = better(m, ArgMax{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)});
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = m.v;
    si[threadIdx.x >> 5] = m.i;
  }
  __syncthreads();
  ArgMax r{sv[0], si[0]};
  for (int w = 1; w < kSampleWarps; ++w) r = better(r, ArgMax{sv[w], si[w]});
  __syncthreads();
  return r;  // same value in every thread
}
__device__ float block_sum(float v, float* sv) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < kSampleWarps; ++w) r += sv[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kSampleThreads) sample_greedy_kernel(SampleArgs a) {
  const int b = blockIdx.x;
  const int slot = a.row_slot[b];
  if (a.slot_done[slot]) return;
  const int state = a.slot_state ? a.slot_state[slot] : -1;
  const float* lg = a.logits + static_cast<size_t>(b) * a.ldl;
  const uint32_t* mask =
      (state >= 0 && a.mask_bits) ? a.mask_bits + static_cast<size_t>(state) * a.mask_words
                                  : nullptr;
  __shared__ float sv[kSampleWarps];
  __shared__ int si[kSampleWarps];
  ArgMax m{-INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < a.vocab; i += kSampleThreads) {
    if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
    m = better(m, ArgMax{lg[i], i});
  }
  m = block_argmax(m, sv, si);
  int tok = m.i;
  if (tok < 0 || tok >= a.vocab) tok = a.eos_id;  // nothing selectable: terminate the row
  float logprob = 0.f;
  if (a.slot_cum_logprob != nullptr) {  // log-softmax of the masked logits at the arg-max
    float se = 0.f;
    for (int i = threadIdx.x; i < a.vocab; i += kSampleThreads) {
      if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
      se += __expf(lg[i] - m.v);
    }
    se = block_sum(se, sv);
    logprob = -__logf(se);
  }
  if (threadIdx.x == 0) finish_row(a, slot, state, tok, logprob);
}

// ---------------------------------------------------------------------------
// temperature / top-k / top-p sampling.  One CTA per row, everything exact on the
// fp32 logits: radix-select on the order-preserving integer image of z = logit / T gives
// the k-th largest value (ties kept); the same descent weighted by exp(z - max) gives the
// smallest value whose upper tail holds top_p of the mass; the token is then drawn by
// inverse CDF over the kept set in vocabulary order with one Philox4x32-10 draw per
// (seed, row, step).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ uint4 philox4x32_10(uint2 key, uint4 c) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return c;
}

__global__ void __launch_bounds__(kSampleThreads) sample_random_kernel(SampleArgs a) {
  const int b = blockIdx.x;
  const int slot = a.row_slot[b];
  if (a.slot_done[slot]) return;
  const int state = a.slot_state ? a.slot_state[slot] : -1;
  const float* lg = a.logits + static_cast<size_t>(b) * a.ldl;
  const uint32_t* mask =
      (state >= 0 && a.mask_bits) ? a.mask_bits + static_cast<size_t>(state) * a.mask_words
                                  : nullptr;
  const float inv_t = 1.0f / a.temperature;
  const int V = a.vocab;
  auto allowed = [&](int i) { return !mask || ((mask[i >> 5] >> (i & 31)) & 1u); };

  __shared__ float sv[kSampleWarps];
  __shared__ int si[kSampleWarps];
  __shared__ uint32_t hist_u[256];
  __shared__ float hist_f[256];
  __shared__ uint32_t sh_prefix, sh_pmask;
  __shared__ int sh_k;
  __shared__ float sh_above;
  __shared__ float wsum[kSampleWarps + 1];
  __shared__ int sh_tok;

  // 1. max of z over the allowed tokens
  ArgMax m{-INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < V; i += kSampleThreads)
    if (allowed(i)) m = better(m, ArgMax{lg[i] * inv_t, i});
  m = block_argmax(m, sv, si);
  if (m.i < 0 || m.i >= V) {  // nothing selectable
    if (threadIdx.x == 0) finish_row(a, slot, state, a.eos_id, 0.f);
    return;
  }
  const float zmax = m.v;

  // 2. top-k threshold (key of the k-th largest z), ties kept
  uint32_t tau = 0;
  if (a.top_k > 0) {
    if (threadIdx.x == 0) sh_prefix = 0, sh_pmask = 0, sh_k = a.top_k;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += kSampleThreads) hist_u[i] = 0;
      __syncthreads();
      const uint32_t prefix = sh_prefix, pmask = sh_pmask;
      for (int i = threadIdx.x; i < V; i += kSampleThreads) {
        if (!allowed(i)) continue;
        const uint32_t key = fkey(lg[i] * inv_t);
        if ((key & pmask) == prefix) atomicAdd(&hist_u[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int cum = 0, sel = 0;
        for (int bin = 255; bin >= 0; --bin) {
          const int c = static_cast<int>(hist_u[bin]);
          if (cum + c >= sh_k) {
            sel = bin;
            break;
          }
          cum += c;
        }
        sh_k -= cum;
        sh_prefix = prefix | (static_cast<uint32_t>(sel) << shift);
        sh_pmask = pmask | (255u << shift);
      }
      __syncthreads();
    }
    tau = sh_prefix;
  }

  // 3. softmax mass of the (top-k filtered) set; top-p threshold on the same ordering
  float W = 0.f;
  for (int i = threadIdx.x; i < V; i += kSampleThreads) {
    if (!allowed(i)) continue;
    const float z = lg[i] * inv_t;
    if (fkey(z) >= tau) W += __expf(z - zmax);
  }
  W = block_sum(W, sv);
  if (a.top_p < 1.0f) {
    const float target = a.top_p * W;
    if (threadIdx.x == 0) sh_prefix = 0, sh_pmask = 0, sh_above = 0.f;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 256; i += kSampleThreads) hist_f[i] = 0.f;
      __syncthreads();
      const uint32_t prefix = sh_prefix, pmask = sh_pmask;
      for (int i = threadIdx.x; i < V; i += kSampleThreads) {
        if (!allowed(i)) continue;
        const float z = lg[i] * inv_t;
        const uint32_t key = fkey(z);
        if (key >= tau && (key & pmask) == prefix)
          atomicAdd(&hist_f[(key >> shift) & 255u], __expf(z - zmax));
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        float cum = sh_above;
        int sel = -1, last = 0;
        for (int bin = 255; bin >= 0; --bin) {
          const float c = hist_f[bin];
          if (c > 0.f) last = bin;
          if (c > 0.f && cum + c >= target) {
            sel = bin;
            break;
          }
          cum += c;
        }
        if (sel < 0) {  // rounding: the tail never quite reached the target -> keep everything
          sel = last;
          cum -= hist_f[last];
        }
        sh_above = cum;
        sh_prefix = prefix | (static_cast<uint32_t>(sel) << shift);
        sh_pmask = pmask | (255u << shift);
      }
      __syncthreads();
    }
    if (sh_prefix > tau) tau = sh_prefix;
    __syncthreads();
  }

  // 4. inverse-CDF draw over the kept set in vocabulary order.  Warp w owns a contiguous
  //    chunk; inside it, (iteration, lane) order == index order.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = ((V + kSampleWarps - 1) / kSampleWarps + 31) & ~31;
  const int c0 = warp * chunk, c1 = min(V, c0 + chunk);
  auto weight = [&](int i) -> float {
    if (i >= c1 || !allowed(i)) return 0.f;
    const float z = lg[i] * inv_t;
    return fkey(z) >= tau ? __expf(z - zmax) : 0.f;
  };
  float part = 0.f;
  for (int i = c0 + lane; i < c1; i += 32) part += weight(i);
  part = warp_sum(part);
  if (lane == 0) wsum[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float run = 0.f;
    for (int w = 0; w < kSampleWarps; ++w) {
      const float t = wsum[w];
      wsum[w] = run;  // exclusive prefix
      run += t;
    }
    wsum[kSampleWarps] = run;
    sh_tok = -1;
  }
  __syncthreads();
  const float total = wsum[kSampleWarps];
  const int row = a.slot_row[slot];
  const uint64_t row_id =
      a.row_ids != nullptr ? static_cast<uint64_t>(a.row_ids[row]) : static_cast<uint64_t>(row);
  const uint64_t ctr_row = a.seed_per_row ? row_id : 0ull;
  const uint4 rnd = philox4x32_10(
      make_uint2(static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32)),
      make_uint4(static_cast<uint32_t>(ctr_row), static_cast<uint32_t>(ctr_row >> 32),
                 static_cast<uint32_t>(a.slot_ngen[slot]), 0u));
  const float u = static_cast<float>(rnd.x >> 8) * (1.0f / 16777216.0f);
  const float target = u * total;
  const float wlo = wsum[warp];
  const float whi = (warp + 1 < kSampleWarps) ? wsum[warp + 1] : total;
  if (target >= wlo && (target < whi || warp == kSampleWarps - 1) && whi > wlo) {
    // this warp's chunk contains the target (the last warp also catches target == total)
    float run = wlo;
    int found = -1, last_kept = -1;
    for (int base = c0; base < c1 && found < 0; base += 32) {
      const float w = weight(base + lane);
      float incl = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, w > 0.f && run + incl > target);
      const unsigned kept = __ballot_sync(0xffffffffu, w > 0.f);
      if (kept) last_kept = base + 31 - __clz(kept);
      if (hit) found = base + __ffs(hit) - 1;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (found < 0) found = last_kept;  // fp rounding at the very end of the chunk
    if (lane == 0 && found >= 0) sh_tok = found;
  }
  __syncthreads();
  int tok = sh_tok;
  if (tok < 0) tok = m.i;  // degenerate rounding: fall back to the mode
  float logprob = 0.f;
  if (a.slot_cum_logprob != nullptr) {
    float se = 0.f;  // log-softmax over all allowed tokens at temperature T
    for (int i = threadIdx.x; i < V; i += kSampleThreads)
      if (allowed(i)) se += __expf(lg[i] * inv_t - zmax);
    se = block_sum(se, sv);
    logprob = (lg[tok] * inv_t - zmax) - __logf(se);
  }
  if (threadIdx.x == 0) finish_row(a, slot, state, tok, logprob);
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
  if (a.temperature > 0.f)
    sample_random_kernel<<<a.B, kSampleThreads, 0, stream>>>(a);
  else
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
