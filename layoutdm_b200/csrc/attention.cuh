// Per-(layout, head) self-attention of the LayoutDM denoiser on tcgen05: O = softmax(Q K^T) V, 125 keys, head_dim 58.
// (nn.MultiheadAttention inside Block._sa_block, T/models/transformer_utils.py:140-142,191-205; no masks.)
//
// Input  qkv [M][1536] 16-bit, per-head column blocks padded 58 -> 64 with zeros:
//        Q_h = cols [h*64, h*64+64), K_h = 512 + ..., V_h = 1024 + ... ; Q is already scaled by 1/sqrt(58).
// Output att [M][512] 16-bit, head h in cols [h*64, h*64+64) (cols 58..63 of every head are exact zeros; the out-projection
//        weight is packed with matching zero columns) = A operand of the out-projection.
//
// Persistent CTAs (two per SM, 160 threads) walk the (layout, head) work items, 128 query rows x 128 keys each:
//   warp 0 (one thread) : TMA loads of the Q / K / V head tiles (128B swizzle) and all tcgen05.mma issue:
//                           S[128x128] = Q K^T      A = Q (K-major), B = K (K-major), 4 MMAs of k=16, fp32 in TMEM cols 0..127
//                           O[128x64]  = P V        A = P (K-major, written by the softmax warps), B = V as loaded
//                                                   ([key][d] rows = MN-major operand), 8 MMAs of k=16, TMEM cols 128..191
//   warps 1..4          : thread = query row.  Exact softmax from TMEM (max pass, then exp2 / sum in registers), the
//                         normalised probabilities are rounded to the operand dtype and written as the P operand tile;
//                         later O is read back from TMEM, packed and TMA-stored.
// The next item's tiles are requested once the PV MMAs have retired (P lives in the Q | K area); the other two CTAs on the SM
// fill the bubbles.
#pragma once
#include "common.cuh"

namespace ldm {

constexpr int kAttThreads = 160;
constexpr int kAttTile = 128 * 128;                 // one 128 x 64 16-bit tile = 16 KB
// smem: Q | K (later overwritten by the 2 k-blocks of P: Q and K are dead once the S MMAs retired) | V | O staging | barriers
// 64 KB + TMEM 128 columns per CTA -> three CTAs per SM
constexpr int kAttOffQ = 0, kAttOffK = kAttTile, kAttOffP = 0, kAttOffV = 2 * kAttTile, kAttOffO = 3 * kAttTile;
constexpr int kAttOffBar = 4 * kAttTile;
constexpr int kAttSmemBytes = 4 * kAttTile + 128 + 1024;   // + barriers + alignment slack
constexpr uint32_t kAttTmemCols = 128;              // S: cols 0..127; O reuses cols 0..63 once the softmax has consumed S

// MN-major (rows = K index, 64 contiguous 16-bit elements = N) operand tile with 128-byte swizzle, 8-row groups 1024 B apart
LDM_DEVINL uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;               // LBO: single 64-element atom along N, unused
  d |= static_cast<uint64_t>(1024 >> 4) << 32;       // SBO: stride between groups of 8 K-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool BF16>
__global__ void __launch_bounds__(kAttThreads, 3)
attention_kernel(const __grid_constant__ CUtensorMap map_qkv /*[M][1536], box 64 x 128*/,
                 const __grid_constant__ CUtensorMap map_att /*[M][512], box 64 x 128*/, int n_valid /*125*/, int n_heads /*8*/,
                 int n_layouts) {
  using O = OpT<BF16>;
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAttOffBar);
  uint64_t* qk_full = bars + 0;   // tx: Q + K tiles
  uint64_t* v_full = bars + 1;    // tx: V tile
  uint64_t* s_full = bars + 2;    // commit: S ready (and Q, K smem free)
  uint64_t* p_ready = bars + 3;   // 128 arrivals: P tile written, S consumed
  uint64_t* o_full = bars + 4;    // commit: O ready (and P, V smem free)
  uint64_t* o_done = bars + 5;    // 128 arrivals: O consumed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_layouts * n_heads;                     // item = layout * n_heads + head

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_att);
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(o_full, 1);
    mbar_init(p_ready, 128); mbar_init(o_done, 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, kAttTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(128, 128, BF16 ? 1 : 0);                 // A, B K-major
      constexpr uint32_t idesc_o = make_idesc_f16(128, 64, BF16 ? 1 : 0) | (1u << 16);     // B (= V) MN-major
      const uint32_t sQ = smem_u32(smem + kAttOffQ), sK = smem_u32(smem + kAttOffK), sV = smem_u32(smem + kAttOffV), sP = smem_u32(smem + kAttOffP);
      auto load_qk = [&](int item) {
        const int h = item % n_heads, row0 = (item / n_heads) * 128;
        mbar_arrive_expect_tx(qk_full, 2 * kAttTile);
        tma_load_2d(smem + kAttOffQ, &map_qkv, qk_full, h * 64, row0);
        tma_load_2d(smem + kAttOffK, &map_qkv, qk_full, n_heads * 64 + h * 64, row0);
      };
      auto load_v = [&](int item) {
        const int h = item % n_heads, row0 = (item / n_heads) * 128;
        mbar_arrive_expect_tx(v_full, kAttTile);
        tma_load_2d(smem + kAttOffV, &map_qkv, v_full, 2 * n_heads * 64 + h * 64, row0);
      };
      const int step = gridDim.x;
      if (static_cast<int>(blockIdx.x) < n_items) { load_qk(blockIdx.x); load_v(blockIdx.x); }
      int hi = 0;
      for (int item = blockIdx.x; item < n_items; item += step, ++hi) {
        const uint32_t ph = hi & 1;
        const bool has_next = item + step < n_items;
        // ---- S = Q K^T ----
        mbar_wait(qk_full, ph);
        tc_fence_after();
        {
          const uint64_t da = make_smem_desc_sw128(sQ), db = make_smem_desc_sw128(sK);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tS, da + 2 * k, db + 2 * k, idesc_s, k != 0);
        }
        umma_commit(s_full);
        // ---- O = P V ----
        mbar_wait(p_ready, ph);                                  // P written (and S fully read)
        mbar_wait(v_full, ph);
        if (hi > 0) mbar_wait(o_done, (hi - 1) & 1);             // previous O has been read out of TMEM
        tc_fence_after();
        {
          const uint64_t da = make_smem_desc_sw128(sP), db = make_smem_desc_mn_sw128(sV);
#pragma unroll
          for (int k = 0; k < 8; ++k)                            // 16 keys per MMA: P advances 32 B inside a k-block / 16 KB
            umma_f16(tO, da + (k >> 2) * (kAttTile >> 4) + 2 * (k & 3), db + k * (2048 >> 4), idesc_o, k != 0);   // V: 16 rows
        }
        umma_commit(o_full);
        mbar_wait(o_full, ph);                                   // P (= the Q | K area) and V are free again
        if (has_next) { load_qk(item + step); load_v(item + step); }
      }
    }
  } else {
    // ===================== softmax / output warps: thread = query row =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                              // row of the 128-row tile
    const uint32_t tl = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t sP = smem_u32(smem + kAttOffP), sO = smem_u32(smem + kAttOffO);
    constexpr float kLog2e = 1.4426950408889634f;
    int hi = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++hi) {
      const uint32_t ph = hi & 1;
      const int h = item % n_heads, row0 = (item / n_heads) * 128;
      mbar_wait(s_full, ph);
      tc_fence_after();
      // pass 1: row maximum over the valid keys
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld<32>(tS + tl + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) if (c * 32 + j < n_valid) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      // pass 2: e = 2^((s - max) log2 e) kept in registers, row sum
      const float mb = mx * kLog2e;
      float e[128];
      float sum = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld<32>(tS + tl + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = (c * 32 + j < n_valid) ? exp2f(fmaf(__uint_as_float(v[j]), kLog2e, -mb)) : 0.0f;
          e[c * 32 + j] = x;
          sum += x;
        }
      }
      const float inv = 1.0f / sum;
      // normalised probabilities, rounded to the operand dtype, as the K-major P tile (2 k-blocks of 64 keys, 128B swizzle)
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const uint4 pk = make_uint4(O::pack(e[8 * c] * inv, e[8 * c + 1] * inv), O::pack(e[8 * c + 2] * inv, e[8 * c + 3] * inv),
                                    O::pack(e[8 * c + 4] * inv, e[8 * c + 5] * inv), O::pack(e[8 * c + 6] * inv, e[8 * c + 7] * inv));
        sts_u4(sP + (c >> 3) * kAttTile + r * 128 + (((c & 7) ^ (r & 7)) << 4), pk);
      }
      fence_proxy_async();                                       // generic-proxy writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- output ----
      mbar_wait(o_full, ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld<32>(tO + tl, o0);
      tmem_ld<32>(tO + tl + 32, o1);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(o_done);                                       // TMEM O may be overwritten by the next head
      if (warp == 1 && lane == 0) bulk_wait_read0();             // the previous head's TMA store has read the staging tile
      named_bar_sync(1, 128);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t* s = c < 4 ? o0 + 8 * c : o1 + 8 * (c - 4);
        const uint4 pk = make_uint4(O::pack(__uint_as_float(s[0]), __uint_as_float(s[1])), O::pack(__uint_as_float(s[2]), __uint_as_float(s[3])),
                                    O::pack(__uint_as_float(s[4]), __uint_as_float(s[5])), O::pack(__uint_as_float(s[6]), __uint_as_float(s[7])));
        sts_u4(sO + r * 128 + ((c ^ (r & 7)) << 4), pk);
      }
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (warp == 1 && lane == 0) { tma_store_2d(&map_att, sO, h * 64, row0); bulk_commit(); }
    }
    if (warp == 1 && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, kAttTmemCols); }
}

}  // namespace ldm
