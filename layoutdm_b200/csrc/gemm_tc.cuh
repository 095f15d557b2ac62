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
// Two kernels:
//   gemm_tc_kernel : N tiled (UMMA_N <= 256), double-buffered accumulators, epilogues
//                    QKV (bias, q-scale) / FF1 (bias, ReLU) / head (fp32 logits)
//   gemm_ln_kernel : the whole d_model = 464 row in one CTA (two MMAs 240 + 224 per k-step), epilogue
//                    bias + residual + LayerNorm (affine or timestep-adaptive): two passes over the TMEM accumulator,
//                    per-thread row statistics combined across the two column halves through shared memory.
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
  static constexpr int kBTileBytes = UMMA_N * kBK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static_assert(kBTileBytes % 1024 == 0, "B tile must keep 1024-B (swizzle atom) alignment");
  static constexpr int kBiasBytes = 1856 * 4;   // the whole bias vector of the layer lives in smem
  static constexpr int kBytes = STAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes;
};

template <int BN_STORE, int UMMA_N, int STAGES, int EPI, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
  using SM = GemmSmem<UMMA_N, STAGES>;
  static_assert(UMMA_N % 16 == 0 && UMMA_N <= 256 && BN_STORE <= UMMA_N, "invalid UMMA shape");
  static_assert(BN_STORE % 8 == 0, "store width");
  constexpr int kAccStride = 256;            // TMEM columns between the two accumulators
  constexpr uint32_t kTmemCols = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * SM::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* sbias = reinterpret_cast<float*>(smem + STAGES * SM::kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int total_tiles = (p.M / kBM) * p.n_tiles;
  for (int i = threadIdx.x; i < p.N; i += kGemmThreads) sbias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.0f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiThreads); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SM::kStageBytes;
          mbar_arrive_expect_tx(&full[stage], SM::kStageBytes);
          tma_load_2d(sa, &map_a, &full[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(sa + kATileBytes, &map_b, &full[stage], kb * kBK, n_blk * BN_STORE);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBM, UMMA_N, BF16 ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + kATileBytes);
          const int nk = min(kBK, p.K - kb * kBK) / kUmmaK;     // K tail: TMA zero-fills, skip the zero k-steps
          for (int k = 0; k < nk; ++k)
            umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);   // +32 B per k-step (>>4 = 2)
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // which half of the tile's columns
    const int row_in_tile = quad * 32 + lane;
    constexpr int kFull = BN_STORE / 32, kRem = BN_STORE % 32;
    constexpr int kSplit = (kFull + 1) / 2;    // half 0: chunks [0, kSplit), half 1: [kSplit, kFull) + remainder
    static_assert(kRem == 0 || kRem == 8 || kRem == 16, "unsupported tile width");
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles, n_blk = tile % p.n_tiles;
      const int n0 = n_blk * BN_STORE;
      const size_t row = static_cast<size_t>(m_blk) * kBM + row_in_tile;
      float tile_scale = 1.0f;
      if constexpr (EPI == EPI_QKV) tile_scale = (n0 < p.qcols) ? p.qscale : 1.0f;   // Q tiles are whole tiles (512 % 256 == 0)
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride;

      auto do_chunk = [&](auto width_tag, int c0) {
        constexpr int W = decltype(width_tag)::value;
        uint32_t r[32];
        tmem_ld<W>(taddr + c0, r);
        float bv[W];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {      // smem broadcast reads overlap the TMEM load
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + n0 + c0 + 4 * j);
          bv[4 * j] = b4.x; bv[4 * j + 1] = b4.y; bv[4 * j + 2] = b4.z; bv[4 * j + 3] = b4.w;
        }
        tmem_wait_ld();
        float v[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          float x = __uint_as_float(r[j]) + bv[j];
          if constexpr (EPI == EPI_QKV) x *= tile_scale;
          if constexpr (EPI == EPI_RELU) x = fmaxf(x, 0.0f);
          v[j] = x;
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
        for (int c = 0; c < kSplit; ++c) do_chunk(std::integral_constant<int, 32>{}, c * 32);
      } else {
#pragma unroll 1
        for (int c = kSplit; c < kFull; ++c) do_chunk(std::integral_constant<int, 32>{}, c * 32);
        if constexpr (kRem == 16) do_chunk(std::integral_constant<int, 16>{}, kFull * 32);
        if constexpr (kRem == 8) do_chunk(std::integral_constant<int, 8>{}, kFull * 32);
      }

      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ---------------------------------------------------------------------------------------------------------
// Full-row GEMM + residual + LayerNorm
// ---------------------------------------------------------------------------------------------------------
constexpr int kD = 464;                     // d_model of the LayoutDM denoiser (512 * 29/32)
constexpr int kLnN1 = 240, kLnN2 = 224;     // the two MMA widths covering 464 columns
constexpr int kLnStages = 3;
constexpr int kLnBTileBytes = kD * kBK * 2;               // 59392 = 58 * 1024
constexpr int kLnStageBytes = kATileBytes + kLnBTileBytes; // 75776
constexpr int kLnStatBytes = 2 * kBM * 8;          // per-row (sum, sumsq) partials of the two column halves
constexpr int kLnSmemBytes = kLnStages * kLnStageBytes + 1024 + 256 + kLnStatBytes;

struct GemmLnParams {
  int M, K;
  const float* bias;       // [464]
  const float* resid;      // fp32 [M][464]
  float* y_out;            // fp32 [M][464] pre-norm sum (the next residual) or nullptr
  const float* ln_scale;   // [464]: gamma (affine) or AdaLN scale (then 1 + scale is applied)
  const float* ln_shift;   // [464]: beta or AdaLN shift
  int adaln;               // 1: out = norm * (1 + scale) + shift
  float* out32;            // fp32 [M][464] normalised output (residual of the next block) or nullptr
  void* out16;             // 16-bit [M][464] normalised output = next GEMM's A operand
};

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmLnParams p) {
  static_assert(kLnBTileBytes % 1024 == 0 && (kLnN1 * kBK * 2) % 1024 == 0, "swizzle atom alignment");
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kLnStages * kLnStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kLnStages;
  uint64_t* tfull = bars + 2 * kLnStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 1);
  float2* sstat = reinterpret_cast<float2*>(smem + kLnStages * kLnStageBytes + 256);   // [2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int total_tiles = p.M / kBM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < kLnStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kLnStageBytes;
          mbar_arrive_expect_tx(&full[stage], kLnStageBytes);
          tma_load_2d(sa, &map_a, &full[stage], kb * kBK, tile * kBM);
          tma_load_2d(sa + kATileBytes, &map_b, &full[stage], kb * kBK, 0);                          // W rows 0..231
          tma_load_2d(sa + kATileBytes + (kD / 2) * kBK * 2, &map_b, &full[stage], kb * kBK, kD / 2);  // rows 232..463
          if (++stage == kLnStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_f16(kBM, kLnN1, BF16 ? 1 : 0);
      constexpr uint32_t idesc2 = make_idesc_f16(kBM, kLnN2, BF16 ? 1 : 0);
      int stage = 0; uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kLnStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db1 = make_smem_desc_sw128(sa + kATileBytes);
          const uint64_t db2 = make_smem_desc_sw128(sa + kATileBytes + kLnN1 * kBK * 2);
          const int nk = min(kBK, p.K - kb * kBK) / kUmmaK;
          for (int k = 0; k < nk; ++k) {
            umma_f16(tmem_base, da + 2 * k, db1 + 2 * k, idesc1, (kb | k) != 0);
            umma_f16(tmem_base + kLnN1, da + 2 * k, db2 + 2 * k, idesc2, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == kLnStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull);
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue: 8 warps = 4 lane quadrants x 2 column halves of 232 =====================
    using O = OpT<BF16>;
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    constexpr int kHalfCols = kD / 2;                 // 232 = 7 * 32 + 8
    const int col0 = half * kHalfCols;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const size_t row = static_cast<size_t>(tile) * kBM + row_in_tile;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col0;
      const float* rrow = p.resid + row * kD + col0;

      // pass 1: y = acc + bias + resid ; y back to TMEM (and to y_out) ; partial row sum / sum of squares
      float sum = 0.0f, sq = 0.0f;
      auto pass1 = [&](auto width_tag, int c0) {
        constexpr int W = decltype(width_tag)::value;
        uint32_t r[32];
        tmem_ld<W>(taddr + c0, r);
        float4 rs[W / 4], bs[W / 4];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          rs[j] = __ldg(reinterpret_cast<const float4*>(rrow + c0) + j);
          bs[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c0) + j);
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          float4 y;
          y.x = __uint_as_float(r[4 * j + 0]) + bs[j].x + rs[j].x;
          y.y = __uint_as_float(r[4 * j + 1]) + bs[j].y + rs[j].y;
          y.z = __uint_as_float(r[4 * j + 2]) + bs[j].z + rs[j].z;
          y.w = __uint_as_float(r[4 * j + 3]) + bs[j].w + rs[j].w;
          sum += (y.x + y.y) + (y.z + y.w);
          sq = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, sq))));
          r[4 * j + 0] = __float_as_uint(y.x); r[4 * j + 1] = __float_as_uint(y.y);
          r[4 * j + 2] = __float_as_uint(y.z); r[4 * j + 3] = __float_as_uint(y.w);
          if (p.y_out != nullptr) reinterpret_cast<float4*>(p.y_out + row * kD + col0 + c0)[j] = y;
        }
        tmem_st<W>(taddr + c0, r);
      };
#pragma unroll 1
      for (int c = 0; c < kHalfCols / 32; ++c) pass1(std::integral_constant<int, 32>{}, c * 32);
      pass1(std::integral_constant<int, 8>{}, (kHalfCols / 32) * 32);
      sstat[half * kBM + row_in_tile] = make_float2(sum, sq);
      tmem_wait_st();
      named_bar_sync(1, kEpiThreads);                  // both halves of every row have published their partials
      const float2 other = sstat[(half ^ 1) * kBM + row_in_tile];
      const float mean = (sum + other.x) * (1.0f / kD);
      const float var = fmaxf((sq + other.y) * (1.0f / kD) - mean * mean, 0.0f);
      const float rstd = 1.0f / sqrtf(var + 1e-5f);

      // pass 2: normalise, scale/shift, store
      const float gadd = p.adaln ? 1.0f : 0.0f;
      auto pass2 = [&](auto width_tag, int c0) {
        constexpr int W = decltype(width_tag)::value;
        uint32_t r[32];
        tmem_ld<W>(taddr + c0, r);
        float4 gs[W / 4], hs[W / 4];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          gs[j] = __ldg(reinterpret_cast<const float4*>(p.ln_scale + col0 + c0) + j);
          hs[j] = __ldg(reinterpret_cast<const float4*>(p.ln_shift + col0 + c0) + j);
        }
        tmem_wait_ld();
        float v[W];
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          v[4 * j + 0] = (__uint_as_float(r[4 * j + 0]) - mean) * rstd * (gs[j].x + gadd) + hs[j].x;
          v[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * (gs[j].y + gadd) + hs[j].y;
          v[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * (gs[j].z + gadd) + hs[j].z;
          v[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * (gs[j].w + gadd) + hs[j].w;
        }
        if (p.out32 != nullptr) {
          float4* dst = reinterpret_cast<float4*>(p.out32 + row * kD + col0 + c0);
#pragma unroll
          for (int j = 0; j < W / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        uint4* dst16 = reinterpret_cast<uint4*>(static_cast<typename O::T*>(p.out16) + row * kD + col0 + c0);
#pragma unroll
        for (int j = 0; j < W / 8; ++j)
          dst16[j] = make_uint4(O::pack(v[8 * j], v[8 * j + 1]), O::pack(v[8 * j + 2], v[8 * j + 3]),
                                O::pack(v[8 * j + 4], v[8 * j + 5]), O::pack(v[8 * j + 6], v[8 * j + 7]));
      };
#pragma unroll 1
      for (int c = 0; c < kHalfCols / 32; ++c) pass2(std::integral_constant<int, 32>{}, c * 32);
      pass2(std::integral_constant<int, 8>{}, (kHalfCols / 32) * 32);

      tc_fence_before();
      mbar_arrive(tempty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace ldm
