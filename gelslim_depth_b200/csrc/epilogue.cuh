// Shared conv epilogue: TMEM accumulator -> scale/shift (folded BatchNorm / bias) -> ReLU -> bf16 ->
// global memory, optional fused 2x2 max-pool (warp shuffles) and optional fused OutConv 1x1 head.
// One thread owns one pixel (= one TMEM lane = one accumulator row).  EIGHT epilogue warps share a tile:
// two per TMEM lane quadrant, taking alternate 32-column units, so that two warps per SM sub-partition
// hide each other's tcgen05.ld / store latency.
//
// Stores: a thread holds 64 contiguous bytes (32 channels) of ITS pixel, so storing straight from registers
// makes every lane of a warp-wide 16-byte store hit a different 128-byte line (32 LSU wavefronts per
// instruction; measured LSU-bound).  Instead each warp transposes through a private 2 KB swizzled smem
// patch (no block-level barrier): afterwards 4 consecutive lanes write one pixel's 64 bytes, i.e. a
// warp-wide store covers 8 pixels x 64 B with full sectors.
#pragma once
#include <cuda_bf16.h>

#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kEpiWarps = 8;
constexpr int kEpiStageBytesPerWarp = 32 * 64;   // 32 pixels x 32 channels bf16


struct EpiPixel {
  __nv_bfloat16* prow;    // &pooled[window][channel base + this lane's 8-channel piece] or null
  __nv_bfloat16* rp[4];   // &out[pixel of warp row 8*i + lane/4][tile channel base] (null: outside the image)
  float* s_stats;         // smem [2][stats_stride] per-CTA partial (sum, sum of squares) of the RAW accumulator, or null
  int stats_ch0;          // absolute channel of accumulator column 0 of this tile
  int stats_stride;       //   (train-mode BatchNorm batch statistics, unet.py:12 in .train()); indexed by absolute channel
  bool valid;             // this lane's pixel lies inside the image (only such pixels enter the statistics)
  bool store_out;         // warp-uniform: the bf16 activation is written
  bool pvalid;            // pooling window inside the pooled image
  bool hx, hy;            // which half this lane keeps in the x / y pooling exchange
  int ypart;              // lane xor mask of the vertical pooling partner (= tile width)
};

// Processes the 32 accumulator columns [c0, c0+32) of this thread's row.
//   sc/sh  : smem scale/shift already offset to the tile's first channel
//   stage  : this warp's private staging patch (shared-window address)
//   hacc/g_head/ncls: fused 1x1 head partial sums / smem weights [ncls][64] (g_head == nullptr: no head)
__device__ __forceinline__ void epilogue_32cols(uint32_t t_row, int c0, const float* sc, const float* sh, int relu,
                                                const EpiPixel& px, uint32_t stage, int lane, float (&hacc)[4],
                                                const float* g_head, int ncls) {
  uint32_t v[32];
  tmem_ld32(t_row + c0, v);
  tmem_ld_wait();
  if (px.s_stats) {
    // Batch statistics (sum, sum of squares of the raw fp32 accumulators) of 32 channels over this warp's 32 pixels, through
    // the warp's 2 KB staging patch (it is free here: the bf16 transposition below runs after it).  Two halves of 16
    // channels: every lane writes 16 accumulators of its pixel as a 64-byte row (chunk j at j ^ ((row >> 1) & 3): the
    // STS.128 of a quarter-warp hit eight different 16-byte bank groups), then lanes 0-15 sum one channel each over the
    // even rows and lanes 16-31 over the odd rows (one 128-byte wavefront per LDS), and one shuffle joins the two.
    // 64 smem wavefronts + 4 shuffles per unit and warp.  The 62-shuffle butterfly this replaces cost ~3.5x that on the
    // one-wavefront-per-clock pipe that also feeds the tensor core's operands (ncu counts shuffles as LSU shared-memory
    // wavefronts): +30 % on a 64-channel train-mode forward conv.
    const int ch = lane & 15, odd = lane >> 4;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t addr = stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
        const int k = 16 * h + 4 * j;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(px.valid ? v[k] : 0u), "r"(px.valid ? v[k + 1] : 0u),
                     "r"(px.valid ? v[k + 2] : 0u), "r"(px.valid ? v[k + 3] : 0u) : "memory");
      }
      __syncwarp();
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int r2 = 0; r2 < 16; ++r2) {
        const uint32_t addr = stage + (2 * r2 + odd) * 64 + (((ch >> 2) ^ (r2 & 3)) << 4) + ((ch & 3) << 2);
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
        a1 += x;
        a2 = fmaf(x, x, a2);
      }
      __syncwarp();
      a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
      a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
      if (odd == h) { s1 = a1; s2 = a2; }        // lane l ends up with channel c0 + l
    }
    atomicAdd(px.s_stats + px.stats_ch0 + c0 + lane, s1);
    atomicAdd(px.s_stats + px.stats_stride + px.stats_ch0 + c0 + lane, s2);
  }
  // per-channel constants are warp-uniform smem reads: fetch them as 128-bit broadcasts (one smem wavefront per
  // 4 channels) -- the smem data pipe is shared with the tensor core's operand reads and is the scarce resource.
  // A broadcast LDS.128 still costs two smem wavefronts, so the constants of a 32-column unit cost as much smem
  // bandwidth as the bf16 transposition below: callers pass sc == nullptr for "scale is 1" (training forward, dgrad) and
  // sh == nullptr for "shift is 0" (dgrad), and those loads disappear.
  float f[32];
  if (sc && sh) {
    const float4* sc4 = reinterpret_cast<const float4*>(sc + c0);
    const float4* sh4 = reinterpret_cast<const float4*>(sh + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = sc4[i], b = sh4[i];
      ffma2(f[4 * i + 0], f[4 * i + 1], __uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]), a.x, a.y, b.x, b.y);
      ffma2(f[4 * i + 2], f[4 * i + 3], __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), a.z, a.w, b.z, b.w);
    }
  } else if (sc) {
    const float4* sc4 = reinterpret_cast<const float4*>(sc + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = sc4[i];
      fmul2(f[4 * i + 0], f[4 * i + 1], __uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]), a.x, a.y);
      fmul2(f[4 * i + 2], f[4 * i + 3], __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), a.z, a.w);
    }
  } else if (sh) {
    const float4* sh4 = reinterpret_cast<const float4*>(sh + c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = sh4[i];
      fadd2(f[4 * i + 0], f[4 * i + 1], __uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1]), b.x, b.y);
      fadd2(f[4 * i + 2], f[4 * i + 3], __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]), b.z, b.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  }
  if (g_head) {
    // the fused 1x1 head consumes the fp32 post-ReLU values (only the network's last conv takes this path)
    if (relu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < ncls) {           // warp-uniform
        const float4* w4 = reinterpret_cast<const float4*>(g_head + k * 64 + c0);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 w = w4[i];
          s = fmaf(f[4 * i + 0], w.x, s);
          s = fmaf(f[4 * i + 1], w.y, s);
          s = fmaf(f[4 * i + 2], w.z, s);
          s = fmaf(f[4 * i + 3], w.w, s);
        }
        hacc[k] += s;
      }
    }
  }
  // ReLU rides on the bf16 conversion (cvt.rn.relu): no FMNMX per element
  uint32_t pk[16];
  if (relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2_relu(f[2 * i], f[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  }
  if (px.store_out) {
    // lane -> row `lane` of the patch; 16-byte chunk j lives at (j ^ ((row >> 1) & 3)): conflict-free both ways
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t addr = stage + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                   "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 8 * i + (lane >> 2), c = lane & 3;
      uint4 val;
      const uint32_t addr = stage + r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(addr));
      if (px.rp[i]) *reinterpret_cast<uint4*>(px.rp[i] + c0 + c * 8) = val;
    }
    __syncwarp();
  }
  if (px.prow) {
    // 2x2 max over lanes {l, l^1, l^ypart, l^1^ypart}; every exchange halves the channels a lane keeps, so the
    // window's 4 lanes end up with 8 distinct channels each (12 shuffles instead of 32, 1 store per lane).
    uint32_t m8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t send = px.hx ? pk[i] : pk[i + 8];
      const uint32_t keep = px.hx ? pk[i + 8] : pk[i];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
      __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&keep), *reinterpret_cast<const __nv_bfloat162*>(&recv));
      m8[i] = *reinterpret_cast<uint32_t*>(&m);
    }
    uint32_t m4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t send = px.hy ? m8[i] : m8[i + 4];
      const uint32_t keep = px.hy ? m8[i + 4] : m8[i];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, px.ypart);
      __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&keep), *reinterpret_cast<const __nv_bfloat162*>(&recv));
      m4[i] = *reinterpret_cast<uint32_t*>(&m);
    }
    if (px.pvalid) *reinterpret_cast<uint4*>(px.prow + c0) = make_uint4(m4[0], m4[1], m4[2], m4[3]);
  }
}

}  // namespace gsd
