// tcgen05 (5th-gen tensor core) GEMMs of the LayoutDM denoiser, C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue).
//
//   A : activations, row-major [M][K] 16-bit (fp16 or bf16), M = 128 * n_layouts (one M-tile = one layout)
//   W : nn.Linear weight, row-major [N][K] 16-bit  (both operands are "K-major" for the MMA)
//
// Warp-specialised persistent kernels, 320 threads:
//   warp 0     : TMA producer  (cp.async.bulk.tensor 2-D tiles, 128-byte swizzle, mbarrier complete_tx)
//   warp 1     : TMEM allocation + single-thread tcgen05.mma issue (fp32 accumulators in TMEM)
//   warps 2..9 : epilogue; warp w owns TMEM lanes 32*(w%4)..+31 (thread = one output row) and one half of the
//                tile's columns (two warps per scheduler, so one warp's TMEM/global latency hides behind the other)
//
// gemm_tc_kernel : N tiled (UMMA_N <= 256), double-buffered TMEM accumulators so the epilogue of tile i overlaps the
//                  main loop of tile i+1; epilogues QKV (bias, q-scale, 16-bit) / FF1 (bias, ReLU, 16-bit) /
//                  fp32 (bias; out-projection, FF2 and the vocabulary head).  The epilogue touches global memory only
//                  with stores: a 1-CTA/SM kernel with ~200 KB of smem has no L1 and only 8 epilogue warps to hide load
//                  latency (round-1a/1b profiles: a residual+LayerNorm epilogue fused here ran 10x slower than the MMAs it
//                  followed), so residual add + LayerNorm live in resid_ln_kernel (embed.cuh), a bandwidth-bound kernel
//                  with full occupancy.
//
// Reference ops replaced: nn.Linear / nn.MultiheadAttention projections / nn.LayerNorm / AdaLayerNorm in
// T/models/transformer_utils.py:79-83,165-210 and T/models/common/nn_lib.py:187-189,235.
#pragma once
#include "common.cuh"

namespace ldm {

constexpr int kBM = 128;       // rows per M tile (= one layout: 125 tokens + 3 pad rows)
constexpr int kBK = 64;        // K elements per smem stage (= 128 B = one swizzle row)
constexpr int kUmmaK = 16;     // K per tcgen05.mma (16-bit operands)
constexpr int kGemmThreads = 320;
constexpr int kEpiThreads = 256;   // warps 2..9
constexpr int kATileBytes = kBM * kBK * 2;   // 16 KB

enum : int { EPI_QKV = 0, EPI_RELU = 1, EPI_F32 = 2 };

struct GemmParams {
  int M, N, K;            // M multiple of 128; N = n_tiles * BN_STORE
  int n_tiles;
  const float* bias;      // [N] or nullptr
  void* out;              // 16-bit [M][ldo] (EPI_QKV / EPI_RELU) or float [M][ldo] (EPI_F32)
  int ldo;
  float qscale;           // EPI_QKV: columns < qcols are scaled by qscale after the bias
  int qcols;
};

template <int UMMA_N, int STAGES>
struct GemmSmem {
  static constexpr int kBHalfBytes = (UMMA_N / 2) * kBK * 2;     // each CTA of the pair holds half of the weight tile
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  static_assert(kBHalfBytes % 1024 == 0, "half weight tile must keep 1024-B (swizzle atom) alignment");
  static constexpr int kBiasBytes = 1856 * 4;   // the whole bias vector of the layer lives in smem
  static constexpr int kXposeBytes = 8 * 4096;  // per-epilogue-warp 32 x 128 B transpose buffer for coalesced stores
  static constexpr int kBytes = STAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes + kXposeBytes;
};

// CTA pairs (thread-block cluster of 2, cta_group::2): the pair computes a 256 x UMMA_N tile per step with ONE
// tcgen05.mma issued by the leader CTA (rank 0): A = 128 rows from each CTA's own smem, B = UMMA_N/2 weight rows from each
// CTA's smem, D = 128 accumulator rows in each CTA's TMEM.  Per CTA and k-block only 16 KB (A) + ~15 KB (half of B) enter
// shared memory for 480 tensor-core cycles: ~65 B/clk, inside what one SM can ingest.  (Round 1c/1d: with 1-CTA 128 x 240
// tiles the 46 KB per k-block needed ~98 B/clk; every k-block took ~870 cycles instead of 480 -- with or without TMA
// multicast, which only saves L2 reads, not the SM's ingest.)
template <int BN_STORE, int UMMA_N, int STAGES, int EPI, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
  using SM = GemmSmem<UMMA_N, STAGES>;
  static_assert(UMMA_N % 16 == 0 && UMMA_N <= 256 && BN_STORE <= UMMA_N, "invalid UMMA shape");
  static_assert(BN_STORE % 8 == 0, "store width");
  constexpr int kAccStride = 256;            // TMEM columns between the two accumulators
  constexpr uint32_t kTmemCols = 512;
  constexpr int kHalfRows = UMMA_N / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * SM::kStageBytes);
  uint64_t* full = bars;                     // leader's copy is the live one: 2 producer arrivals + both CTAs' TMA bytes
  uint64_t* empty = bars + STAGES;           // per CTA: released by the leader's multicast tcgen05.commit
  uint64_t* tfull = bars + 2 * STAGES;       // per CTA: accumulator ready (multicast commit)
  uint64_t* tempty = bars + 2 * STAGES + 2;  // leader's copy: 16 warp arrivals (8 epilogue warps x 2 CTAs)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* sbias = reinterpret_cast<float*>(smem + STAGES * SM::kStageBytes + 256);
  static_assert((2 * STAGES + 4) * 8 + 4 <= 256, "barrier block overflow");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const uint32_t cta_rank = cluster_ctarank();               // 0 = leader
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int total_tiles = (p.M / (2 * kBM)) * p.n_tiles;     // tiles per pair: (256-row block, N-tile)
  for (int i = threadIdx.x; i < p.N; i += kGemmThreads) sbias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.0f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 16); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // both CTAs' barriers and TMEM are ready
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (one thread in each CTA) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int m_blk = 2 * (tile / p.n_tiles) + static_cast<int>(cta_rank), n_blk = tile % p.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SM::kStageBytes;
          const uint32_t lead_full = mapa_shared(smem_u32(&full[stage]), 0);
          mbar_arrive_expect_tx_cluster(lead_full, SM::kStageBytes);
          tma_load_2d_2cta(sa, &map_a, lead_full, kb * kBK, m_blk * kBM);
          tma_load_2d_2cta(sa + kATileBytes, &map_b, lead_full, kb * kBK, n_blk * BN_STORE + static_cast<int>(cta_rank) * kHalfRows);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * kBM, UMMA_N, BF16 ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);              // both CTAs drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);                    // both CTAs' A and B halves have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + kATileBytes);
          const int nk = min(kBK, p.K - kb * kBK) / kUmmaK;     // K tail: TMA zero-fills, skip the zero k-steps
          for (int k = 0; k < nk; ++k)
            umma_f16_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);   // +32 B per k-step (>>4 = 2)
          umma_commit_2cta_mc(&empty[stage], static_cast<uint16_t>(0b11));       // free the stage in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta_mc(&tfull[acc], static_cast<uint16_t>(0b11));           // accumulator ready in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // which half of the tile's columns
    const int row_in_tile = quad * 32 + lane;
    const uint32_t sbias_addr = smem_u32(sbias);
    const uint32_t xbuf = smem_u32(smem + STAGES * SM::kStageBytes + 256 + SM::kBiasBytes) + (warp - 2) * 4096;
    constexpr int kFull = BN_STORE / 32, kRem = BN_STORE % 32;
    constexpr int kSplit = (kFull + 1) / 2;    // half 0: chunks [0, kSplit), half 1: [kSplit, kFull) + remainder
    static_assert(kRem == 0 || kRem == 8 || kRem == 16, "unsupported tile width");
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int m_blk = 2 * (tile / p.n_tiles) + static_cast<int>(cta_rank), n_blk = tile % p.n_tiles;
      const int n0 = n_blk * BN_STORE;
      const size_t row = static_cast<size_t>(m_blk) * kBM + row_in_tile;
      float tile_scale = 1.0f;
      if constexpr (EPI == EPI_QKV) tile_scale = (n0 < p.qcols) ? p.qscale : 1.0f;   // Q tiles are whole tiles (512 % 256 == 0)
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;

      // 32 accumulator columns -> registers -> (+bias, activation) -> this warp's smem transpose buffer -> row-contiguous
      // 16-B global stores: every store instruction writes whole 64/128-byte row segments instead of 32 scattered 16-B
      // pieces (round-1f profile: 38 % of epilogue stall cycles were LG-queue throttling on the scattered stores).
      auto do_chunk32 = [&](int c0) {
        uint32_t r[32];
        tmem_ld<32>(taddr + c0, r);
        float bv[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {            // LDS broadcast reads overlap the TMEM load
          const float4 b4 = lds_f4(sbias_addr + (n0 + c0 + 4 * j) * 4);
          bv[4 * j] = b4.x; bv[4 * j + 1] = b4.y; bv[4 * j + 2] = b4.z; bv[4 * j + 3] = b4.w;
        }
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]) + bv[j];
          if constexpr (EPI == EPI_QKV) x *= tile_scale;
          if constexpr (EPI == EPI_RELU) x = fmaxf(x, 0.0f);
          v[j] = x;
        }
        __syncwarp();                            // previous chunk's read-back of the buffer is complete
        if constexpr (EPI == EPI_F32) {
          // row = lane, 8 x 16-B pieces per 128-B row, piece c at (c ^ (lane & 7))
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts_u4(xbuf + lane * 128 + ((c ^ (lane & 7)) << 4),
                   make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3])));
          __syncwarp();
          float* obase = static_cast<float*>(p.out) + (static_cast<size_t>(m_blk) * kBM + quad * 32) * p.ldo + n0 + c0;
#pragma unroll
          for (int it = 0; it < 8; ++it) {       // each instruction: 4 rows x 128 B
            const int rr = it * 4 + (lane >> 3), c = lane & 7;
            const uint4 d = lds_u4(xbuf + rr * 128 + ((c ^ (rr & 7)) << 4));
            *reinterpret_cast<uint4*>(obase + static_cast<size_t>(rr) * p.ldo + c * 4) = d;
          }
        } else {
          using O = OpT<BF16>;
          // row = lane, 4 x 16-B pieces per 64-B row (two rows per 128 B), piece c at (c ^ ((lane >> 1) & 3))
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts_u4(xbuf + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4),
                   make_uint4(O::pack(v[8 * c], v[8 * c + 1]), O::pack(v[8 * c + 2], v[8 * c + 3]),
                              O::pack(v[8 * c + 4], v[8 * c + 5]), O::pack(v[8 * c + 6], v[8 * c + 7])));
          __syncwarp();
          typename O::T* obase = static_cast<typename O::T*>(p.out) + (static_cast<size_t>(m_blk) * kBM + quad * 32) * p.ldo + n0 + c0;
#pragma unroll
          for (int it = 0; it < 4; ++it) {       // each instruction: 8 rows x 64 B
            const int rr = it * 8 + (lane >> 2), c = lane & 3;
            const uint4 d = lds_u4(xbuf + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
            *reinterpret_cast<uint4*>(obase + static_cast<size_t>(rr) * p.ldo + c * 8) = d;
          }
        }
      };
      // remainder columns (8 or 16): direct row stores
      auto do_rem = [&](auto width_tag, int c0) {
        constexpr int W = decltype(width_tag)::value;
        uint32_t r[32];
        tmem_ld<W>(taddr + c0, r);
        tmem_wait_ld();
        float v[W];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          const float4 b4 = lds_f4(sbias_addr + (n0 + c0 + 4 * j) * 4);
          v[4 * j] = __uint_as_float(r[4 * j]) + b4.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
          v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
          if constexpr (EPI == EPI_QKV) v[j] *= tile_scale;
          if constexpr (EPI == EPI_RELU) v[j] = fmaxf(v[j], 0.0f);
        }
        if constexpr (EPI == EPI_F32) {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row * p.ldo + n0 + c0);
#pragma unroll
          for (int j = 0; j < W / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          using O = OpT<BF16>;
          uint4* dst = reinterpret_cast<uint4*>(static_cast<typename O::T*>(p.out) + row * p.ldo + n0 + c0);
#pragma unroll
          for (int j = 0; j < W / 8; ++j)
            dst[j] = make_uint4(O::pack(v[8 * j], v[8 * j + 1]), O::pack(v[8 * j + 2], v[8 * j + 3]),
                                O::pack(v[8 * j + 4], v[8 * j + 5]), O::pack(v[8 * j + 6], v[8 * j + 7]));
        }
      };
      if (half == 0) {
#pragma unroll 1
        for (int c = 0; c < kSplit; ++c) do_chunk32(c * 32);
      } else {
#pragma unroll 1
        for (int c = kSplit; c < kFull; ++c) do_chunk32(c * 32);
        if constexpr (kRem == 16) do_rem(std::integral_constant<int, 16>{}, kFull * 32);
        if constexpr (kRem == 8) do_rem(std::integral_constant<int, 8>{}, kFull * 32);
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));   // leader's barrier: 8 warps x 2 CTAs
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // nobody exits while the peer may still arrive here / read our smem
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2cta(tmem_base, kTmemCols); }
}

}  // namespace ldm
