// Per-(layout, head) self-attention of the LayoutDM denoiser: O = softmax(Q K^T) V, S = 125 keys, head_dim 58.
// (nn.MultiheadAttention inside Block._sa_block, T/models/transformer_utils.py:140-142,191-205; no masks.)
//
// Input  qkv [M][1536] 16-bit, per-head column blocks padded 58 -> 64 with zeros:
//        Q_h = cols [h*64, h*64+64), K_h = 512 + ..., V_h = 1024 + ... ; Q is already scaled by 1/sqrt(58).
// Output att [M][464] 16-bit, heads concatenated compactly (col = h*58 + j) = A operand of the out-projection.
//
// One CTA = 4 heads of one layout, head after head (the next head's Q/K/V tiles stream in through cp.async while the
// current one is computed); per head: 128 query rows (125 valid) x 128 keys (125 valid, the rest masked to -inf).
// 8 warps x 16 query rows; the whole score row lives in registers, so the softmax is exact (max, exp, sum,
// normalise) before the probabilities are rounded to the operand dtype -- the same rounding points as the
// oracle's same-rounding mode.  Contractions use warp-level mma.sync m16n8k16 (fp32 accumulate): attention
// is 1.07 % of the denoiser FLOPs; the 98.9 % in the linear layers run on tcgen05.
#pragma once
#include "common.cuh"

namespace ldm {

constexpr int kAttThreads = 256;
constexpr int kAttHeadsPerCta = 4;               // a CTA walks 4 heads of one layout with double-buffered staging
constexpr int kAttTileBytes = 3 * 128 * 128;     // Q, K, V tiles of one head: 128 rows x 64 x 2 B each
constexpr int kAttSmemBytes = 2 * kAttTileBytes; // two heads in flight

template <bool BF16>
LDM_DEVINL void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
LDM_DEVINL void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
LDM_DEVINL void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// byte offset of 16-byte chunk `c` (0..7) of row `r` in a [128][64 x 16-bit] tile, XOR-swizzled so that the
// eight rows an ldmatrix phase touches fall into distinct banks
LDM_DEVINL uint32_t att_off(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

template <bool BF16>
__global__ void __launch_bounds__(kAttThreads, 2)
attention_kernel(const void* __restrict__ qkv_, void* __restrict__ att_, int n_valid /*125*/, int head_dim /*58*/, int n_heads /*8*/) {
  using O = OpT<BF16>;
  using T = typename O::T;
  extern __shared__ __align__(128) uint8_t att_smem[];
  const T* qkv = static_cast<const T*>(qkv_);
  T* att = static_cast<T*>(att_);
  const int groups = n_heads / kAttHeadsPerCta;
  const int layout = blockIdx.x / groups, h0 = (blockIdx.x % groups) * kAttHeadsPerCta;
  const int ldq = 3 * n_heads * 64;
  const size_t row0 = static_cast<size_t>(layout) * 128;
  const uint32_t sbase = smem_u32(att_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  const int ldo = n_heads * head_dim;

  // stage the Q, K, V tiles of head h into buffer `buf` (cp.async, 16 B per thread per op)
  auto stage = [&](int h, int buf) {
    for (int i = threadIdx.x; i < 3 * 128 * 8; i += kAttThreads) {
      const int mat = i / (128 * 8), r = (i / 8) % 128, c = i % 8;
      const T* src = qkv + (row0 + r) * ldq + mat * (n_heads * 64) + h * 64 + c * 8;
      const uint32_t dst = sbase + buf * kAttTileBytes + mat * (128 * 128) + att_off(r, c);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  stage(h0, 0);
  for (int hi = 0; hi < kAttHeadsPerCta; ++hi) {
    const int h = h0 + hi, buf = hi & 1;
    if (hi + 1 < kAttHeadsPerCta) {
      stage(h + 1, buf ^ 1);                                  // next head streams in while this one is computed
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sQ = sbase + buf * kAttTileBytes, sK = sQ + 128 * 128, sV = sQ + 2 * 128 * 128;

    // ---- S = Q K^T : 16 x 128 per warp ----
    uint32_t qf[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int r = m0 + (lane & 15);          // lanes 0-15: rows 0-15 (k chunk 2kt); lanes 16-31: same rows, chunk 2kt+1
      const int c = 2 * kt + (lane >> 4);
      ldsm_x4(qf[kt], sQ + att_off(r, c));
    }
    float s[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f; }
#pragma unroll
    for (int np = 0; np < 8; ++np) {            // pairs of 8-key tiles
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        // matrices: (keys np*16+0..7, k chunk 2kt), (same keys, chunk 2kt+1), (keys +8.., chunk 2kt), (keys +8.., chunk 2kt+1)
        const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = 2 * kt + ((lane >> 3) & 1);
        uint32_t kf[4];
        ldsm_x4(kf, sK + att_off(r, c));
        mma16816<BF16>(s[2 * np], qf[kt], kf[0], kf[1]);
        mma16816<BF16>(s[2 * np + 1], qf[kt], kf[2], kf[3]);
      }
    }

    // ---- softmax over the n_valid keys (rows g and g+8 of this warp's 16) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int key = nt * 8 + 2 * t + j;
        if (key >= n_valid) { s[nt][j] = -INFINITY; s[nt][2 + j] = -INFINITY; }
        mx0 = fmaxf(mx0, s[nt][j]);
        mx1 = fmaxf(mx1, s[nt][2 + j]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    // exp(x - m) = 2^((x - m) * log2 e): one FFMA + MUFU.EX2; the probabilities are rounded to 16 bits right after, so the
    // 2-ulp approximation is far below the operand rounding
    constexpr float kLog2e = 1.4426950408889634f;
    const float mb0 = mx0 * kLog2e, mb1 = mx1 * kLog2e;
    float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        s[nt][j] = exp2f(fmaf(s[nt][j], kLog2e, -mb0)); sum0 += s[nt][j];
        s[nt][2 + j] = exp2f(fmaf(s[nt][2 + j], kLog2e, -mb1)); sum1 += s[nt][2 + j];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

    // ---- O = P V : 16 x 64 per warp ----
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.0f; }
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {            // 16 keys per step
      uint32_t pf[4];
      pf[0] = O::pack(s[2 * kt][0] * inv0, s[2 * kt][1] * inv0);
      pf[1] = O::pack(s[2 * kt][2] * inv1, s[2 * kt][3] * inv1);
      pf[2] = O::pack(s[2 * kt + 1][0] * inv0, s[2 * kt + 1][1] * inv0);
      pf[3] = O::pack(s[2 * kt + 1][2] * inv1, s[2 * kt + 1][3] * inv1);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {          // pairs of 8-wide d tiles
        // trans matrices: (keys kt*16+0..7, d chunk 2dp), (keys +8.., chunk 2dp), (keys 0..7, chunk 2dp+1), (keys +8.., chunk 2dp+1)
        const int r = kt * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int c = 2 * dp + (lane >> 4);
        uint32_t vf[4];
        ldsm_x4_trans(vf, sV + att_off(r, c));
        mma16816<BF16>(o[2 * dp], pf, vf[0], vf[1]);
        mma16816<BF16>(o[2 * dp + 1], pf, vf[2], vf[3]);
      }
    }

    // ---- store (heads compact: col = h*head_dim + d, d < head_dim) ----
    T* out0 = att + (row0 + m0 + g) * ldo + h * head_dim;
    T* out1 = out0 + 8 * static_cast<size_t>(ldo);
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      const int d = nd * 8 + 2 * t;
      if (d < head_dim) {                        // head_dim is even: a pair never straddles the boundary
        *reinterpret_cast<uint32_t*>(out0 + d) = O::pack(o[nd][0], o[nd][1]);
        *reinterpret_cast<uint32_t*>(out1 + d) = O::pack(o[nd][2], o[nd][3]);
      }
    }
    __syncthreads();                             // everyone is done with `buf` before the head after next overwrites it
  }
}

}  // namespace ldm
