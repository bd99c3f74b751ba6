// sutro_b200 — K1: bf16 GEMM  D[M,N] = A[M,K] · W[N,K]^T  on tcgen05 tensor cores.
//
// Every projection of the transformer (QKV, O, gate/up, down, lm_head) has this
// shape: activations row-major [tokens, K], weights row-major [N, K] — both
// K-major, which is the native operand layout of tcgen05.mma with 128-byte
// swizzled shared-memory tiles.
//
// Kernel anatomy (persistent, warp-specialised, one CTA per SM):
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D loads of A (128x64) and
//             W (BLOCK_N x 64) tiles into a kStages-deep shared-memory ring.
//   warp 1    MMA issuer: one lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16)
//             four times per stage, accumulating in TMEM; tcgen05.commit frees
//             the stage / publishes the accumulator through mbarriers.
//   warp 2    TMEM allocator (2 accumulator stages so the epilogue of tile i
//             overlaps the main loop of tile i+1).
//   warps 4-7 epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused
//             epilogue -> 64-byte-per-thread global stores.
//
// Fused epilogues (rounding points mirror the bf16 oracle, oracle/model_ref.py):
//   STORE_BF16     D = bf16(acc)
//   RESIDUAL_BF16  D = bf16(bf16(acc) + R)           (O-proj / down-proj + residual)
//   SWIGLU_BF16    D[:, j] = bf16(silu(bf16(acc[2j])) * bf16(acc[2j+1]))
//                  (gate/up weights interleaved row-wise; output has N/2 columns)
//   STORE_F32      D = acc                             (lm_head logits)
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace sb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kAccStages = 2;
constexpr int kEpiWarp0 = 4;
// 4 epilogue warps (one per TMEM lane quarter); the fused QKV epilogue does an order of
// magnitude more math per element, so it gets two warps per quarter, each owning half of
// every head (dims [32p,32p+32) and their +64 RoPE partners).
template <int EPI>
constexpr int epi_warps() { return EPI == EPI_QKV_ROPE ? 8 : 4; }
template <int EPI>
constexpr int gemm_threads() { return 128 + 32 * epi_warps<EPI>(); }

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = kAccStages * BLOCK_N;  // 512 / 256 / 128: powers of two
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +align slack
};

// Tile rasterisation: consecutive tile ids sweep all N tiles of a group of GM row-tiles
// before moving on, so the CTAs running concurrently share a [GM*128, K] slab of A and a
// ~148/GM-tile slab of W that both fit in the 126 MB L2.  (M-fastest order re-read A from
// HBM once per N tile: 5x the algorithmic traffic on the down projection, ncu r01.)
SB_DEVICE void tile_to_mn(int GM, int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  const int group_tiles = GM * num_n;
  const int g = tile / group_tiles;
  const int first_m = g * GM;
  const int gm = min(GM, num_m - first_m);
  const int within = tile - g * group_tiles;
  m_blk = first_m + within % gm;
  n_blk = within / gm;
}

// One thread's 32 consecutive accumulator columns of one row -> fused epilogue -> global.
template <int EPI>
SB_DEVICE void epilogue_store(const uint32_t (&v)[32], void* __restrict__ d_out,
                              const __nv_bfloat16* __restrict__ resid, int row, int col0, int ldd) {
  if constexpr (EPI == EPI_STORE_F32) {
    float* dp = reinterpret_cast<float*>(d_out) + static_cast<size_t>(row) * ldd + col0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      st_v4(dp + 4 * i, make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  } else if constexpr (EPI == EPI_SWIGLU_BF16) {
    __nv_bfloat16* dp =
        reinterpret_cast<__nv_bfloat16*>(d_out) + static_cast<size_t>(row) * ldd + (col0 >> 1);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float r2[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float g = bf16_round(__uint_as_float(v[4 * i + 2 * h]));
        const float u = bf16_round(__uint_as_float(v[4 * i + 2 * h + 1]));
        const float s = bf16_round(g / (1.0f + __expf(-g)));
        r2[h] = s * u;
      }
      o[i] = pack_bf16x2(r2[0], r2[1]);
    }
    st_v4(dp, make_uint4(o[0], o[1], o[2], o[3]));
    st_v4(dp + 8, make_uint4(o[4], o[5], o[6], o[7]));
  } else {
    __nv_bfloat16* dp =
        reinterpret_cast<__nv_bfloat16*>(d_out) + static_cast<size_t>(row) * ldd + col0;
    uint32_t o[16];
    if constexpr (EPI == EPI_RESIDUAL_BF16) {
      const __nv_bfloat16* rp = resid + static_cast<size_t>(row) * ldd + col0;
      uint4 rv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) rv[i] = *reinterpret_cast<const uint4*>(rp + 8 * i);
      const uint32_t* ru = reinterpret_cast<const uint32_t*>(rv);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 r = unpack_bf16x2(ru[i]);
        o[i] = pack_bf16x2(bf16_round(__uint_as_float(v[2 * i])) + r.x,
                           bf16_round(__uint_as_float(v[2 * i + 1])) + r.y);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        o[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st_v4(dp + 8 * i, make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
  }
}

// ---------------------------------------------------------------------------
// EPI_QKV_ROPE: one thread owns one token row and, per 128-column head, applies exactly
// what rope_kv_kernel (norm_rope.cu) does to the stored bf16 tensor — bf16 rounding of the
// projection, optional per-head RMSNorm, rotate-half RoPE with bf16 op-by-op rounding —
// then writes q heads to the qkv buffer and k/v heads straight into the swizzled KV page (and,
// unswizzled, into their columns of the qkv buffer for the dense prefill attention).
// The whole head lives in this thread, so the reduction needs no shuffles.
// ---------------------------------------------------------------------------
struct QkvRowMeta {
  int pos, r;
  size_t page;
  uint4 cs[4], sn[4];  // this row's cos/sin for the warp's 32 dims, fetched during the mainloop
};

SB_DEVICE void qkv_head_epilogue(uint32_t tmem_head, const QkvEpiArgs& ea, const QkvRowMeta& rm,
                                 bool row_ok, int row, int head, __nv_bfloat16* __restrict__ qkv_out,
                                 int ldd, int part) {
  // 128 fp32 accumulators -> 64 packed bf16x2 (this is the rounding of the linear output)
  // This warp owns dims [32*part, 32*part+32) and [64+32*part, ...): TMEM column chunks
  // `part` and `part+2`.  The other two chunks are read only for the sum of squares.
  const bool is_q = head < ea.hq;
  const bool is_k = !is_q && head < ea.hq + ea.hkv;
  const __nv_bfloat16* nw =
      static_cast<const __nv_bfloat16*>(is_q ? ea.q_norm_w : (is_k ? ea.k_norm_w : nullptr));
  uint32_t pk[32];  // [0,16): low-half dims, [16,32): their +64 partners (bf16x2 packed)
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const bool mine = (c & 1) == part;
    if (!mine && nw == nullptr) continue;  // warp-uniform: nothing needed from this chunk
    uint32_t v[32];
    tmem_ld_32x32(tmem_head + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t u = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      const float2 f = unpack_bf16x2(u);  // the bf16-rounded projection output
      ss += f.x * f.x + f.y * f.y;
      if (mine) pk[(c >> 1) * 16 + i] = u;
    }
  }
  if (!row_ok) return;
  const float rstd = nw != nullptr ? rsqrtf(ss / static_cast<float>(kHeadDim) + ea.eps) : 1.f;
  __nv_bfloat16* dst_lo;
  int swz = 0;
  if (is_q) {
    dst_lo = qkv_out + static_cast<size_t>(row) * ldd + head * kHeadDim;
  } else {
    const int kvh = is_k ? (head - ea.hq) : (head - ea.hq - ea.hkv);
    dst_lo = static_cast<__nv_bfloat16*>(ea.kv_layer) +
             (rm.page * ea.hkv + kvh) * (2 * kTileElems) + (is_k ? 0 : kTileElems) +
             rm.r * kHeadDim;
    swz = rm.r & 7;
  }
  // Packed bf16x2 arithmetic: a product of two bf16 values is exact in fp32, so HMUL2.BF16
  // (one rounding) equals "fp32 multiply, round to bf16" — the rounding points of the unfused
  // kernel (norm_rope.cu) and of the oracle; sums are HADD2.BF16 (exact sum, one rounding).
  // Two elements per instruction and no unpack/repack: the epilogue is a third of its former
  // instruction count, which is what keeps the fused QKV GEMM from being epilogue-bound.
  // explicit .rn forms: ptxas must not contract mul + add/sub into one fused (singly rounded) FMA
  auto mul2 = [](uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  };
  auto add2 = [](uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  };
  auto sub2 = [](uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
  };
#pragma unroll
  for (int gg = 0; gg < 4; ++gg) {  // dims 8g..8g+7 of the low half and their +64 partners
    const int g = 4 * part + gg;
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lo[j] = pk[4 * gg + j];
      hi[j] = pk[16 + 4 * gg + j];
    }
    if (nw != nullptr) {
      const uint4 wl4 = *reinterpret_cast<const uint4*>(nw + 8 * g);
      const uint4 wh4 = *reinterpret_cast<const uint4*>(nw + 64 + 8 * g);
      const uint32_t wl[4] = {wl4.x, wl4.y, wl4.z, wl4.w};
      const uint32_t wh[4] = {wh4.x, wh4.y, wh4.z, wh4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // w * bf16(x * rstd): rstd is fp32, so that product is fp32
        const float2 a = unpack_bf16x2(lo[j]), b = unpack_bf16x2(hi[j]);
        lo[j] = mul2(wl[j], pack_bf16x2(a.x * rstd, a.y * rstd));
        hi[j] = mul2(wh[j], pack_bf16x2(b.x * rstd, b.y * rstd));
      }
    }
    if (is_q || is_k) {
      // out[i] = x[i]*cos - x[i+64]*sin ; out[i+64] = x[i+64]*cos + x[i]*sin (bf16 op by op)
      const uint4 c4 = rm.cs[gg], s4 = rm.sn[gg];
      const uint32_t cu[4] = {c4.x, c4.y, c4.z, c4.w};
      const uint32_t su[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x = lo[j], y = hi[j];
        lo[j] = sub2(mul2(x, cu[j]), mul2(y, su[j]));
        hi[j] = add2(mul2(y, cu[j]), mul2(x, su[j]));
      }
    }
    const uint4 olo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    const uint4 ohi = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    st_v4(dst_lo + ((g ^ swz) << 3), olo);
    st_v4(dst_lo + (((g + 8) ^ swz) << 3), ohi);
    if (!is_q && ea.write_dense) {
      // dense copy next to q: the tcgen05 prefill attention reads the new tokens' K/V from the
      // qkv buffer through TMA (attn_prefill_tc.cu); decode reads the paged copy above
      __nv_bfloat16* dense = qkv_out + static_cast<size_t>(row) * ldd + head * kHeadDim;
      st_v4(dense + (g << 3), olo);
      st_v4(dense + ((g + 8) << 3), ohi);
    }
  }
}

SB_DEVICE void qkv_row_meta(const QkvEpiArgs& ea, int row, bool row_ok, int part, QkvRowMeta& rm) {
  rm.pos = 0, rm.r = 0, rm.page = 0;
  if (row_ok) {
    rm.pos = ea.tok_pos[row];
    const int slot = ea.tok_slot[row];
    rm.page = static_cast<size_t>(
        ea.page_table[static_cast<size_t>(slot) * ea.max_pages + rm.pos / kPageTokens]);
    rm.r = rm.pos % kPageTokens;
  }
  const uint4* cosr = reinterpret_cast<const uint4*>(
      static_cast<const __nv_bfloat16*>(ea.cos_tab) + static_cast<size_t>(rm.pos) * 64);
  const uint4* sinr = reinterpret_cast<const uint4*>(
      static_cast<const __nv_bfloat16*>(ea.sin_tab) + static_cast<size_t>(rm.pos) * 64);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    rm.cs[g] = cosr[4 * part + g];
    rm.sn[g] = sinr[4 * part + g];
  }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a,
                    const __grid_constant__ CUtensorMap tm_b, void* __restrict__ d_out,
                    const __nv_bfloat16* __restrict__ resid, int M, int N, int K, int ldd,
                    int ea_gm, const QkvEpiArgs ea) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = bars + 2 * Cfg::kStages + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 32 * epi_warps<EPI>());
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + kBlockM - 1) / kBlockM;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_k = K / kBlockK;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_to_mn(ea_gm, tile, num_m, num_n, m_blk, n_blk);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_long(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb_ = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
          tma_load_2d(sa, &tm_a, fb, kb * kBlockK, m_blk * kBlockM);
          tma_load_2d(sb_, &tm_b, fb, kb * kBlockK, n_blk * BLOCK_N);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
    const uint32_t a_lo0 = umma_desc_lo_k_sw128(smem_u32(smem));   // stage 0's A tile
    const uint32_t empty0 = smem_u32(&empty_bar[0]);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        if (lane == 0) {
          static_assert(kBlockK / kUmmaK == 4, "tc_mma_kblock4 issues four K=16 steps");
          const uint32_t a_lo = a_lo0 + stage * (Cfg::kStageBytes >> 4);
          tc_mma_kblock4(tmem_d, a_lo, a_lo + (Cfg::kABytes >> 4), idesc, kb != 0,
                         empty0 + stage * 8);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      // the accumulator is complete when every MMA issued so far has retired
      if (lane == 0) tc_commit(smem_u32(&tfull_bar[acc]));
      __syncwarp();
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue =====
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int part = (warp - kEpiWarp0) >> 2;  // 0, or 1 for the second warp of a quarter
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_to_mn(ea_gm, tile, num_m, num_n, m_blk, n_blk);
      const int row = m_blk * kBlockM + q * 32 + lane;
      const bool row_ok = row < M;
      QkvRowMeta rm;
      if constexpr (EPI == EPI_QKV_ROPE) qkv_row_meta(ea, row, row_ok, part, rm);  // overlaps the mainloop
      mbar_wait_long(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      if constexpr (EPI == EPI_QKV_ROPE) {
        static_assert(BLOCK_N % 128 == 0 || EPI != EPI_QKV_ROPE, "fused QKV needs whole heads");
#pragma unroll 1
        for (int hh = 0; hh < BLOCK_N / 128; ++hh) {
          const int head = (n_blk * BLOCK_N) / kHeadDim + hh;
          qkv_head_epilogue(
              tmem_base + acc * BLOCK_N + hh * 128 + (static_cast<uint32_t>(q * 32) << 16), ea, rm,
              row_ok && head * kHeadDim < N, row, head, reinterpret_cast<__nv_bfloat16*>(d_out),
              ldd, part);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(
              tmem_base + acc * BLOCK_N + c * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
          tmem_ld_wait();
          const int col0 = n_blk * BLOCK_N + c * 32;
          if (!row_ok || col0 >= N) continue;
          epilogue_store<EPI>(v, d_out, resid, row, col0, ldd);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a 2-CTA cluster computes a 256x256 tile.
// Each CTA stages its own 128 rows of A and its own 128 of the tile's 256 W rows
// (so per-SM shared-memory fill traffic is 32 KiB per k-block instead of 48 KiB and
// the tensor core reads half of B from the peer SM); the leader CTA's MMA warp issues
// tcgen05.mma.cta_group::2 with M=256, N=256 and both CTAs' TMEM receive their 128
// accumulator rows.  Barrier protocol:
//   full[s]   leader only; both CTAs' TMA loads complete_tx on the leader's barrier
//   empty[s]  per CTA; freed by a multicast tcgen05.commit
//   tfull[a]  per CTA; multicast commit after the last k-block of a tile
//   tempty[a] leader only; every epilogue warp of both CTAs arrives (count 8)
// ---------------------------------------------------------------------------
constexpr int k2HalfBytes = 128 * kBlockK * 2;      // 16 KiB: 128 rows x 64 bf16
constexpr int k2StageBytes = 2 * k2HalfBytes;       // A half-tile + B half-tile per CTA
constexpr int k2BlockN = 256;
constexpr int k2SmemBytes(int stages) { return stages * k2StageBytes + 256 + 1024; }

// KSUB 64-element K spans share one pipeline stage (one full / one empty barrier): with
// KSUB = 2 the issuing thread and the producer pay the barrier round trip once per 128 K
// elements (8 MMAs).  k2Stages counts stages, so shared memory holds k2Stages * KSUB spans.
template <int EPI, int k2Stages, int KSUB = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads<EPI>(), 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a,
                     const __grid_constant__ CUtensorMap tm_b, void* __restrict__ d_out,
                     const __nv_bfloat16* __restrict__ resid, int M, int N, int K, int ldd,
                     int ea_gm, const QkvEpiArgs ea) {
  extern __shared__ uint8_t smem_raw[];
  // both CTAs of the pair must compute the same offsets: the dynamic smem base is the
  // same in every CTA of a kernel, so the alignment fix-up below is identical too.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int kStageBytes = KSUB * k2StageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2Stages * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + k2Stages;
  uint64_t* tfull_bar = bars + 2 * k2Stages;
  uint64_t* tempty_bar = bars + 2 * k2Stages + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * k2Stages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < k2Stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 2 * epi_warps<EPI>());  // epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cta2(smem_u32(tmem_slot), 2 * k2BlockN);
    tmem_relinquish_cta2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + 255) / 256;
  const int num_n = (N + k2BlockN - 1) / k2BlockN;
  const int num_tiles = num_m * num_n;
  const int num_k = K / (kBlockK * KSUB);
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_to_mn(ea_gm, tile, num_m, num_n, m_blk, n_blk);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_long(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb_leader = smem_u32(&full_bar[stage]) & kPeerBitMask;
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          if (leader) mbar_arrive_expect_tx(smem_u32(&full_bar[stage]), 2 * kStageBytes);
#pragma unroll
          for (int sub = 0; sub < KSUB; ++sub) {
            const int k0 = (kb * KSUB + sub) * kBlockK;
            tma_load_2d_cta2(sa + sub * k2StageBytes, &tm_a, fb_leader, k0,
                             m_blk * 256 + rank * 128);
            tma_load_2d_cta2(sa + sub * k2StageBytes + k2HalfBytes, &tm_b, fb_leader, k0,
                             n_blk * k2BlockN + rank * 128);
          }
          if (++stage == k2Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, k2BlockN);
      const uint32_t a_lo0 = umma_desc_lo_k_sw128(smem_u32(smem));   // stage 0's A tile
      const uint32_t empty0 = smem_u32(&empty_bar[0]);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * k2BlockN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          if (lane == 0) {
            static_assert(kBlockK / kUmmaK == 4, "tc_mma_kblock4 issues four K=16 steps");
            uint32_t a_lo = a_lo0 + stage * (kStageBytes >> 4);
#pragma unroll
            for (int sub = 0; sub + 1 < KSUB; ++sub, a_lo += k2StageBytes >> 4)
              tc_mma_kblock4_cta2_nocommit(tmem_d, a_lo, a_lo + (k2HalfBytes >> 4), idesc,
                                           (kb | sub) != 0);
            tc_mma_kblock4_cta2(tmem_d, a_lo, a_lo + (k2HalfBytes >> 4), idesc,
                                (kb != 0) || KSUB > 1, empty0 + stage * 8, 0x3);
          }
          __syncwarp();
          if (++stage == k2Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // the accumulator is complete when every MMA issued so far has retired
        if (lane == 0) tc_commit_cta2_mc(smem_u32(&tfull_bar[acc]), 0x3);
        __syncwarp();
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue (both CTAs: each drains its own 128 accumulator rows) =====
    const int q = warp & 3;
    const int part = (warp - kEpiWarp0) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_to_mn(ea_gm, tile, num_m, num_n, m_blk, n_blk);
      const int row = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32 + lane;
      const bool row_ok = row < M;
      QkvRowMeta rm;
      if constexpr (EPI == EPI_QKV_ROPE) qkv_row_meta(ea, row, row_ok, part, rm);  // overlaps the mainloop
      mbar_wait_long(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      if constexpr (EPI == EPI_QKV_ROPE) {
#pragma unroll 1
        for (int hh = 0; hh < k2BlockN / 128; ++hh) {
          const int head = (n_blk * k2BlockN) / kHeadDim + hh;
          qkv_head_epilogue(
              tmem_base + acc * k2BlockN + hh * 128 + (static_cast<uint32_t>(q * 32) << 16), ea,
              rm, row_ok && head * kHeadDim < N, row, head,
              reinterpret_cast<__nv_bfloat16*>(d_out), ldd, part);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < k2BlockN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(
              tmem_base + acc * k2BlockN + c * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
          tmem_ld_wait();
          const int col0 = n_blk * k2BlockN + c * 32;
          if (!row_ok || col0 >= N) continue;
          epilogue_store<EPI>(v, d_out, resid, row, col0, ldd);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty_bar[acc]) & kPeerBitMask);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cta2(tmem_base, 2 * k2BlockN);
  }
}

// ---------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int rows, cols, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) ^
         (static_cast<size_t>(k.cols) << 20) ^ (static_cast<size_t>(k.box_rows) << 50);
    return h;
  }
};

// Row-major bf16 [rows, cols] tensor, box = [box_rows, 64 cols], 128B swizzle.
int make_tmap(const void* ptr, int rows, int cols, int box_rows, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: %d (ptr=%p rows=%d cols=%d box_rows=%d)",
                   static_cast<int>(r), ptr, rows, cols, box_rows);
    return -1;
  }
  std::lock_guard<std::mutex> g(mu);
  cache.emplace(key, *out);
  return 0;
}

}  // namespace

// Generic bf16 tiled tensor map with 128-byte swizzle (rank 2..5); strides are in bytes for
// dims 1..rank-1 (dim 0 is contiguous).  Used by the attention kernels.
int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable");
    return -1;
  }
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) st[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                   const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(rank %d) failed: %d (ptr=%p dims=%llu,%llu box=%u,%u)",
                   rank, static_cast<int>(r), ptr, (unsigned long long)dims[0],
                   (unsigned long long)dims[1], box[0], box[1]);
    return -1;
  }
  return 0;
}

namespace {

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// Row-tiles per raster group.  A weight matrix that fits L2 with room to spare (<= 64 MB:
// QKV, Wo, down) stays resident whatever the order, so narrow groups win — the CTAs of a wave
// then share few A row-tiles and walk W together (measured at M=32768: GM 1-4 is 2-4 % faster
// than wide groups, tools/gemm_raster_scan.py).  A larger W (gate/up, the LM head) is streamed:
// there the group takes as many row-tiles as keep its A slab (rows x K bf16) within ~48 MB of
// the 126 MB L2, so A is read from HBM once and W once per group.
int raster_group(int tile_rows, int N, int K) {
  static const int forced = [] {   // SB200_GEMM_GM: experiment knob (tools/gemm_raster_scan.py)
    const char* e = getenv("SB200_GEMM_GM");
    return e ? atoi(e) : 0;
  }();
  if (forced > 0) return forced;
  if (static_cast<long>(N) * K * 2 <= (64L << 20)) return 2;
  const long slab = 48L << 20;
  long gm = slab / (static_cast<long>(tile_rows) * K * 2);
  if (gm < 2) gm = 2;
  if (gm > 64) gm = 64;
  return static_cast<int>(gm);
}

template <int BLOCK_N, int EPI>
int launch_cfg(const void* a, int a_rows, const void* w, void* d, const void* resid, int M, int N,
               int K, int ldd, cudaStream_t stream, const QkvEpiArgs& ea) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap tm_a, tm_b;
  if (make_tmap(a, a_rows, K, kBlockM, &tm_a)) return -1;
  if (make_tmap(w, N, K, BLOCK_N, &tm_b)) return -1;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, EPI>;
  SB_SET_MAX_SMEM(kern, Cfg::kSmemBytes);
  const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, gemm_threads<EPI>(), Cfg::kSmemBytes, stream>>>(
      tm_a, tm_b, d, reinterpret_cast<const __nv_bfloat16*>(resid), M, N, K, ldd,
      raster_group(kBlockM, N, K), ea);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int EPI, int STAGES, int KSUB = 1>
int launch_cta2(const void* a, int a_rows, const void* w, void* d, const void* resid, int M, int N,
                int K, int ldd, cudaStream_t stream, const QkvEpiArgs& ea) {
  CUtensorMap tm_a, tm_b;
  if (make_tmap(a, a_rows, K, 128, &tm_a)) return -1;
  if (make_tmap(w, N, K, 128, &tm_b)) return -1;
  auto kern = gemm2_bf16_tn_kernel<EPI, STAGES, KSUB>;
  SB_SET_MAX_SMEM(kern, k2SmemBytes(STAGES * KSUB));
  const int tiles = ((M + 255) / 256) * ((N + k2BlockN - 1) / k2BlockN);
  const int clusters = std::min(tiles, num_sms() / 2);
  kern<<<2 * clusters, gemm_threads<EPI>(), k2SmemBytes(STAGES * KSUB), stream>>>(
      tm_a, tm_b, d, reinterpret_cast<const __nv_bfloat16*>(resid), M, N, K, ldd,
      raster_group(256, N, K), ea);
  SB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int EPI>
int launch_epi(int block_n, const void* a, int a_rows, const void* w, void* d, const void* resid,
               int M, int N, int K, int ldd, cudaStream_t stream, const QkvEpiArgs& ea) {
  switch (block_n) {
    case 512: {  // CTA-pair kernel: 256x256 tile per 2-CTA cluster
      // K = 128 per pipeline stage (3 stages of two 64-element spans) when K allows it: half
      // the barrier round trips per MMA for the issuing thread.  SB200_GEMM_KSUB=1 keeps the
      // 7 x 64 ring.  The K order of the accumulation is the same either way.
      static const int ksub = [] {
        const char* e = getenv("SB200_GEMM_KSUB");
        return e ? atoi(e) : 1;   // measured: 3 x 128 trails 7 x 64 by 1-2 % (shallower ring)
      }();
      if (ksub == 2 && K % (2 * kBlockK) == 0)
        return launch_cta2<EPI, 3, 2>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
      return launch_cta2<EPI, 7>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    }
    case 516:  // experiment knobs (tools/gemm_bench.py): the 7 x 64 ring, a shallower ring
      return launch_cta2<EPI, 7>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    case 514:
      return launch_cta2<EPI, 4>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    case 64:
      if constexpr (EPI == EPI_QKV_ROPE) {
        set_last_error("gemm_bf16_tn: the fused QKV epilogue needs block_n >= 128");
        return -1;
      } else {
        return launch_cfg<64, EPI>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
      }
    case 128:
      return launch_cfg<128, EPI>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    default:
      return launch_cfg<256, EPI>(a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
  }
}

}  // namespace

int gemm_pick_block_n(int M, int N) {
  // Cost = waves x (tile area / efficiency); pick the cheapest configuration.
  //   512: CTA-pair 256x256 per cluster (74 clusters)   64/128/256: one CTA, 128 x BLOCK_N
  // Efficiencies are relative tensor-pipe rates measured for the tile shapes (the 1-CTA
  // shapes are shared-memory-bandwidth limited: TMA fill + MMA operand reads).
  const int sms = num_sms();
  double best = 1e30;
  int pick = 128;
  struct Cand { int id, tm, tn, units; double eff; };
  const Cand cands[] = {{512, 256, 256, sms / 2, 2.0 * 1.00},
                        {256, 128, 256, sms, 0.82},
                        {128, 128, 128, sms, 0.70},
                        {64, 128, 64, sms, 0.45}};
  for (const Cand& c : cands) {
    const long tiles = static_cast<long>((M + c.tm - 1) / c.tm) * ((N + c.tn - 1) / c.tn);
    const long waves = (tiles + c.units - 1) / c.units;
    const double cost = static_cast<double>(waves) * c.tm * c.tn / c.eff;
    if (cost < best) {
      best = cost;
      pick = c.id;
    }
  }
  return pick;
}

int gemm_bf16_tn(const void* a, int a_rows, const void* w, void* d, const void* resid, int M,
                 int N, int K, int ldd, int epilogue, int block_n, cudaStream_t stream,
                 const QkvEpiArgs* qkv_args) {
  if (M <= 0) return 0;
  const QkvEpiArgs ea = qkv_args ? *qkv_args : QkvEpiArgs{};
  if (K % kBlockK != 0 || N % 32 != 0 || a_rows < M) {
    set_last_error("gemm_bf16_tn: unsupported shape M=%d N=%d K=%d a_rows=%d", M, N, K, a_rows);
    return -1;
  }
  if (block_n == 0) {
    block_n = gemm_pick_block_n(M, N);
    // The wide gate/up projection (N = 2*d_ff) also runs on the CTA pair: once the issue loop
    // stopped being the limiter the pair leads the one-CTA 128x256 tile on that shape too
    // (1411 vs 1352 TFLOP/s sustained at M = 32768, tools/gemm_sustained.py).  Every tile
    // shape accumulates in the same order, so this choice changes no result.
    // SB200_GEMM_SWIGLU_BN=256 selects the one-CTA tile for comparison.
    static const int swiglu_bn = [] {
      const char* e = getenv("SB200_GEMM_SWIGLU_BN");
      return e ? atoi(e) : 512;
    }();
    if (epilogue == EPI_SWIGLU_BF16 && block_n == 512 && N >= 8192 &&
        (swiglu_bn == 256 || swiglu_bn == 512))
      block_n = swiglu_bn;
  }
  if (block_n != 64 && block_n != 128 && block_n != 256 && block_n != 512 && block_n != 514 &&
      block_n != 516) {
    set_last_error("gemm_bf16_tn: block_n must be 64/128/256/512(CTA pair), got %d", block_n);
    return -1;
  }
  switch (epilogue) {
    case EPI_STORE_BF16:
      return launch_epi<EPI_STORE_BF16>(block_n, a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    case EPI_RESIDUAL_BF16:
      if (!resid) {
        set_last_error("gemm_bf16_tn: residual epilogue without residual pointer");
        return -1;
      }
      return launch_epi<EPI_RESIDUAL_BF16>(block_n, a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    case EPI_SWIGLU_BF16:
      return launch_epi<EPI_SWIGLU_BF16>(block_n, a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    case EPI_STORE_F32:
      return launch_epi<EPI_STORE_F32>(block_n, a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    case EPI_QKV_ROPE:
      if (!qkv_args || N % kHeadDim != 0 || N != (ea.hq + 2 * ea.hkv) * kHeadDim) {
        set_last_error("gemm_bf16_tn: fused QKV epilogue needs its operands and N = heads*128");
        return -1;
      }
      if (block_n == 64) block_n = 128;
      return launch_epi<EPI_QKV_ROPE>(block_n, a, a_rows, w, d, resid, M, N, K, ldd, stream, ea);
    default:
      set_last_error("gemm_bf16_tn: unknown epilogue %d", epilogue);
      return -1;
  }
}

}  // namespace sb
