// Per-channel additive constant (folded BatchNorm shift / bias) applied by the TENSOR CORE instead of the epilogue.
//
// Why: the conv epilogue reads its per-channel constants as warp-uniform 128-bit shared-memory loads.  ncu on B200
// (profiles/r1_ncu_full_convs_final.md) shows the 64-channel layers bound by the one-wavefront-per-clock smem data pipe
// (tensor-core operand reads + epilogue LSU traffic ~ 100 % of the cycles), and ~60 % of the epilogue's LSU wavefronts
// are those constant loads (scale + shift: 32 LDS.128 per 32-column unit and warp).  With the BatchNorm scale folded
// into the bf16 weights at pack time (w' = w * gamma / sqrt(var + eps), plan.cu) the remaining constant is additive,
// and an additive per-column constant is one more rank-1 update of the accumulator:
//     D[128 x N] = ones[128 x 16] * biasrows[N x 16]^T      (one K = 16 UMMA, accumulate = 0, first instruction of a tile)
// where ones[:, 0] = ones[:, 1] = 1 and biasrows[n] = (hi, lo, 0, ...) is the bf16 split of the fp32 constant
// (hi = bf16(b), lo = bf16(b - hi): the sum carries 16 mantissa bits; products with 1.0 are exact in the fp32
// accumulator).  The epilogue then needs NO per-channel constant at all.
// Operand layout: 32-byte rows, SWIZZLE_32B (the first layer's layout: row r at r*32, the two 16-byte halves swapped when
// bit 2 of r is set), descriptors with SBO = 8 rows; both buffers are built by generic stores once per CTA.
#pragma once
#include <cuda_bf16.h>

#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kBiasOnesBytes = 128 * 32;
// descriptor high word for contiguous 32-byte rows: SBO = 8 rows, version 1, SWIZZLE_32B
constexpr uint32_t kBiasDescHi = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);

__device__ __forceinline__ void bias_st16(uint32_t addr, uint32_t w0) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

// s_bias (1 KB aligned): [ones: 128 rows][bias rows: `rows` rows]; bias points at this CTA's first channel.
// Every thread of the CTA calls it (tid / nthreads); the caller issues fence.proxy.async + a CTA barrier afterwards.
__device__ __forceinline__ void bias_mma_fill(uint32_t s_bias, const float* __restrict__ bias, int rows, int tid, int nthreads) {
  for (int r = tid; r < 128 + rows; r += nthreads) {
    uint32_t w0;
    if (r < 128) {
      w0 = 0x3F803F80u;                                   // (1.0, 1.0) bf16
    } else {
      const float b = __ldg(bias + (r - 128));
      const __nv_bfloat16 hi = __float2bfloat16_rn(b);
      const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
      w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    const uint32_t row = s_bias + r * 32;
    const uint32_t swz = ((r >> 2) & 1) << 4;             // rows 128.. continue the same 256-byte swizzle period (4096 % 256 == 0)
    bias_st16(row + swz, w0);
    bias_st16(row + (swz ^ 16u), 0u);
  }
}

// one elected lane: D[tmem] = ones * biasrows^T (overwrites the accumulator)
template <bool CTA2>
__device__ __forceinline__ void bias_mma_issue(uint32_t d_tmem, uint32_t s_bias, uint32_t idesc) {
  const uint32_t a_lo = ((s_bias & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo = (((s_bias + kBiasOnesBytes) & 0x3FFFFu) >> 4) | (1u << 16);
  if (CTA2) umma_bf16_lohi_2sm(d_tmem, a_lo, kBiasDescHi, b_lo, kBiasDescHi, idesc, 0u);
  else umma_bf16_lohi(d_tmem, a_lo, kBiasDescHi, b_lo, kBiasDescHi, idesc, 0u);
}

}  // namespace gsd
