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

enum : int { EPI_QKV = 0, EPI_RELU = 1, EPI_F32 = 2, EPI_LN = 3 };

struct GemmParams {
  int M, N, K;            // M multiple of 128; N = n_tiles * BN_STORE
  int n_tiles;
  const float* bias;      // [N] or nullptr
  void* out;              // 16-bit [M][ldo] (EPI_QKV / EPI_RELU) or float [M][ldo] (EPI_F32)
  int ldo;
  float qscale;           // EPI_QKV: columns < qcols are scaled by qscale after the bias
  int qcols;
  // EPI_LN (N = 464 = 2 tiles of 232): y = acc + bias + resid ; out = LayerNorm(y) * gamma + beta  (gamma = 1 + scale_t for AdaLN)
  const float* resid;     // fp32 [M][N] residual stream
  float* y_out;           // fp32 [M][N] pre-norm sum (the next residual) or nullptr
  const float* ln_scale;  // [N]
  const float* ln_shift;  // [N]
  int adaln;
  float* out32;           // fp32 [M][N] normalised output (next residual, AdaLN case) or nullptr
  void* out16;            // 16-bit [M][N] normalised output = next GEMM's A operand
};

template <int UMMA_N, int STAGES>
struct GemmSmem {
  static constexpr int kBHalfBytes = (UMMA_N / 2) * kBK * 2;     // each CTA of the pair holds half of the weight tile
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  static_assert(kBHalfBytes % 1024 == 0, "half weight tile must keep 1024-B (swizzle atom) alignment");
  static constexpr int kBiasBytes = 1856 * 4;   // the whole bias vector of the layer lives in smem
  static constexpr int kXposeBytes = 8 * 4096;  // per-epilogue-warp 32 x 128 B transpose buffer for coalesced loads/stores
  static constexpr int kStatBytes = 2 * kBM * 8;  // EPI_LN: per-row (sum, sumsq) partials of the two column halves
  static constexpr int kBytes = STAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes + kXposeBytes + kStatBytes;
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
  const int n_super = p.M / (2 * kBM);                       // 256-row blocks; a pair walks all N tiles of a block in a row
  for (int i = threadIdx.x; i < p.N; i += kGemmThreads) {
    sbias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.0f;
    if constexpr (EPI == EPI_LN) {
      sbias[p.N + i] = __ldg(p.ln_scale + i) + (p.adaln ? 1.0f : 0.0f);
      sbias[2 * p.N + i] = __ldg(p.ln_shift + i);
    }
  }

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
      for (int sup = pair; sup < n_super; sup += n_pairs)
      for (int n_blk = 0; n_blk < p.n_tiles; ++n_blk) {
        const int m_blk = 2 * sup + static_cast<int>(cta_rank);
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
      for (int sup = pair; sup < n_super; sup += n_pairs)
      for (int n_blk = 0; n_blk < p.n_tiles; ++n_blk) {
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
    if constexpr (EPI != EPI_LN) {
    for (int sup = pair; sup < n_super; sup += n_pairs)
    for (int n_blk = 0; n_blk < p.n_tiles; ++n_blk) {
      const int m_blk = 2 * sup + static_cast<int>(cta_rank);
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
      } else {
      // ============ fused residual + LayerNorm epilogue (out-projection / FF2) ============
      // One CTA owns 128 complete rows (both 232-column tiles of a 256-row block run back to back on this pair).
      //   phase A (per tile, overlaps the other tile's MMAs): y = acc + bias + resid -> back into TMEM (+ y_out), row sum / sumsq
      //   phase B (after both tiles): LayerNorm from TMEM, 16-bit (+ fp32) outputs; then the accumulators are released.
      // resid is read and every output is written through this warp's smem transpose buffer, i.e. as whole 64/128-byte
      // row segments per instruction.
      using O = OpT<BF16>;
      static_assert(EPI != EPI_LN || (BN_STORE == 232 && kRem == 8), "LN epilogue is laid out for 2 x 232 columns");
      float2* sstat = reinterpret_cast<float2*>(smem + STAGES * SM::kStageBytes + 256 + SM::kBiasBytes + SM::kXposeBytes);
      const uint32_t sgamma_addr = sbias_addr + p.N * 4, sbeta_addr = sbias_addr + 2 * p.N * 4;
      const int c_begin = half == 0 ? 0 : kSplit, c_end = half == 0 ? kSplit : kFull;   // 32-col chunks of each tile
      for (int sup = pair; sup < n_super; sup += n_pairs) {
        const int m_blk = 2 * sup + static_cast<int>(cta_rank);
        const size_t row = static_cast<size_t>(m_blk) * kBM + row_in_tile;
        const size_t wrow0 = static_cast<size_t>(m_blk) * kBM + quad * 32;      // first row of this warp
        float sum = 0.0f, sq = 0.0f;
        // pull this warp's residual rows (32 rows x 2 x 116/128 columns) from HBM into L2 while the MMAs run: the
        // dependent loads of phase A then see L2 latency instead of DRAM latency
        {
          const int cols_per_tile = (c_end - c_begin) * 32;             // 128 (half 0) / 96 (half 1); remainder rides along
          const int lines_per_row = (cols_per_tile * 4 + 127) / 128 + (half == 1 ? 1 : 0);
          for (int i = lane; i < 2 * 32 * lines_per_row; i += 32) {
            const int t = i / (32 * lines_per_row), rem = i % (32 * lines_per_row);
            const int rr = rem / lines_per_row, ln = rem % lines_per_row;
            const float* a = p.resid + (wrow0 + rr) * p.N + t * BN_STORE + c_begin * 32 + ln * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
        // ---------------- phase A ----------------
        for (int n_blk = 0; n_blk < 2; ++n_blk) {
          const int n0 = n_blk * BN_STORE;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + n_blk * kAccStride;
          // prefetch the first chunk's residual rows while the MMAs are still running
          uint4 q[8];
          auto load_resid = [&](int c0) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + (lane >> 3), c = lane & 7;
              q[it] = __ldg(reinterpret_cast<const uint4*>(p.resid + (wrow0 + rr) * p.N + n0 + c0 + c * 4));
            }
          };
          load_resid(c_begin * 32);
          mbar_wait(&tfull[n_blk], acc_phase);
          tc_fence_after();
#pragma unroll 1
          for (int c = c_begin; c < c_end; ++c) {
            const int c0 = c * 32;
            uint32_t r[32];
            tmem_ld<32>(taddr + c0, r);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + (lane >> 3), cc = lane & 7;
              sts_u4(xbuf + rr * 128 + ((cc ^ (rr & 7)) << 4), q[it]);
            }
            __syncwarp();
            float y[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 rs = lds_f4(xbuf + lane * 128 + ((j ^ (lane & 7)) << 4));
              const float4 b4 = lds_f4(sbias_addr + (n0 + c0 + 4 * j) * 4);
              y[4 * j] = rs.x + b4.x; y[4 * j + 1] = rs.y + b4.y; y[4 * j + 2] = rs.z + b4.z; y[4 * j + 3] = rs.w + b4.w;
            }
            if (c + 1 < c_end) load_resid(c0 + 32);          // next chunk's residual: in flight during the math below
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              y[j] += __uint_as_float(r[j]);
              sum += y[j];
              sq = fmaf(y[j], y[j], sq);
              r[j] = __float_as_uint(y[j]);
            }
            tmem_st<32>(taddr + c0, r);
            if (p.y_out != nullptr) {
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts_u4(xbuf + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
              __syncwarp();
              float* obase = p.y_out + wrow0 * p.N + n0 + c0;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3), cc = lane & 7;
                *reinterpret_cast<uint4*>(obase + static_cast<size_t>(rr) * p.N + cc * 4) = lds_u4(xbuf + rr * 128 + ((cc ^ (rr & 7)) << 4));
              }
            }
          }
          if (half == 1) {                                    // the 8 remainder columns (224..231) of the tile: direct
            const int c0 = kFull * 32;
            uint32_t r[32];
            tmem_ld<8>(taddr + c0, r);
            float y[8];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float4 rs = __ldg(reinterpret_cast<const float4*>(p.resid + row * p.N + n0 + c0) + j);
              const float4 b4 = lds_f4(sbias_addr + (n0 + c0 + 4 * j) * 4);
              y[4 * j] = rs.x + b4.x; y[4 * j + 1] = rs.y + b4.y; y[4 * j + 2] = rs.z + b4.z; y[4 * j + 3] = rs.w + b4.w;
            }
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              y[j] += __uint_as_float(r[j]);
              sum += y[j];
              sq = fmaf(y[j], y[j], sq);
              r[j] = __float_as_uint(y[j]);
            }
            tmem_st<8>(taddr + c0, r);
            if (p.y_out != nullptr) {
              float4* dst = reinterpret_cast<float4*>(p.y_out + row * p.N + n0 + c0);
              dst[0] = make_float4(y[0], y[1], y[2], y[3]);
              dst[1] = make_float4(y[4], y[5], y[6], y[7]);
            }
          }
        }
        // ---------------- row statistics across the two column halves ----------------
        sstat[half * kBM + row_in_tile] = make_float2(sum, sq);
        tmem_wait_st();
        named_bar_sync(1, kEpiThreads);
        const float2 other = sstat[(half ^ 1) * kBM + row_in_tile];
        const float mean = (sum + other.x) * (1.0f / (2 * BN_STORE));
        const float var = fmaxf((sq + other.y) * (1.0f / (2 * BN_STORE)) - mean * mean, 0.0f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        // ---------------- phase B ----------------
        for (int n_blk = 0; n_blk < 2; ++n_blk) {
          const int n0 = n_blk * BN_STORE;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + n_blk * kAccStride;
          auto norm4 = [&](const uint32_t* rr4, int col, float* o) {
            const float4 g4 = lds_f4(sgamma_addr + col * 4), h4 = lds_f4(sbeta_addr + col * 4);
            o[0] = (__uint_as_float(rr4[0]) - mean) * rstd * g4.x + h4.x;
            o[1] = (__uint_as_float(rr4[1]) - mean) * rstd * g4.y + h4.y;
            o[2] = (__uint_as_float(rr4[2]) - mean) * rstd * g4.z + h4.z;
            o[3] = (__uint_as_float(rr4[3]) - mean) * rstd * g4.w + h4.w;
          };
#pragma unroll 1
          for (int c = c_begin; c < c_end; ++c) {
            const int c0 = c * 32;
            uint32_t r[32];
            tmem_ld<32>(taddr + c0, r);
            tmem_wait_ld();
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) norm4(r + 4 * j, n0 + c0 + 4 * j, v + 4 * j);
            __syncwarp();
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              sts_u4(xbuf + lane * 64 + ((cc ^ ((lane >> 1) & 3)) << 4),
                     make_uint4(O::pack(v[8 * cc], v[8 * cc + 1]), O::pack(v[8 * cc + 2], v[8 * cc + 3]),
                                O::pack(v[8 * cc + 4], v[8 * cc + 5]), O::pack(v[8 * cc + 6], v[8 * cc + 7])));
            __syncwarp();
            typename O::T* o16 = static_cast<typename O::T*>(p.out16) + wrow0 * p.N + n0 + c0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rr = it * 8 + (lane >> 2), cc = lane & 3;
              *reinterpret_cast<uint4*>(o16 + static_cast<size_t>(rr) * p.N + cc * 8) = lds_u4(xbuf + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
            }
            if (p.out32 != nullptr) {
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts_u4(xbuf + lane * 128 + ((j ^ (lane & 7)) << 4),
                       make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
              __syncwarp();
              float* o32 = p.out32 + wrow0 * p.N + n0 + c0;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3), cc = lane & 7;
                *reinterpret_cast<uint4*>(o32 + static_cast<size_t>(rr) * p.N + cc * 4) = lds_u4(xbuf + rr * 128 + ((cc ^ (rr & 7)) << 4));
              }
            }
          }
          if (half == 1) {
            const int c0 = kFull * 32;
            uint32_t r[32];
            tmem_ld<8>(taddr + c0, r);
            tmem_wait_ld();
            float v[8];
            norm4(r, n0 + c0, v);
            norm4(r + 4, n0 + c0 + 4, v + 4);
            *reinterpret_cast<uint4*>(static_cast<typename O::T*>(p.out16) + row * p.N + n0 + c0) =
                make_uint4(O::pack(v[0], v[1]), O::pack(v[2], v[3]), O::pack(v[4], v[5]), O::pack(v[6], v[7]));
            if (p.out32 != nullptr) {
              float4* dst = reinterpret_cast<float4*>(p.out32 + row * p.N + n0 + c0);
              dst[0] = make_float4(v[0], v[1], v[2], v[3]);
              dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[n_blk]), 0));   // release this accumulator
        }
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // nobody exits while the peer may still arrive here / read our smem
  if (warp == 1) { tc_fence_after(); tmem_dealloc_2cta(tmem_base, kTmemCols); }
}

}  // namespace ldm
