// Weight gradient of a 3x3 convolution as a tcgen05 GEMM whose K dimension is the PIXELS:
//
//   dW[co][tap][ci] = sum_{b,y,x} dZ[b,y,x,co] * X[b, y+dy, x+dx, ci]
//
// Both operands are NHWC, i.e. channel-contiguous = "MN-major" for this GEMM (M = co, N = ci, K = pixel), which
// tcgen05 reads directly (instruction-descriptor a_major = b_major = 1): no transposition pass.
//   A = dZ tile  : one TMA box (64 co, 8, 16) per 64-channel block = 128 pixels x 128 B, pixels are the K rows
//   B = X halo   : the same (16+2)x(8+2) halo box the forward kernel uses; the tap shift is again only a shifted
//                  descriptor start ((2k+dy)*10+dx rows) with the 8-pixel K atoms 1280 bytes apart
// One UMMA (K = 16) covers two image rows of the 16x8 tile.  The three dx taps of one filter row are ONE instruction
// with N = 3 x 64: their B operands are the same halo shifted by one pixel each, i.e. three MN blocks exactly 128 bytes
// apart (the descriptor's leading-byte-offset).  That makes the instruction tensor-bound (A 32 + B 48 smem wavefronts
// per 96 tensor cycles) where three N = 64 instructions were smem-bound (3 x 48 per 96).  A CTA owns one
// (128 co) x (64 ci) x (filter rows {0,1} or {2}) block of dW in TMEM (2 x 192 or 192 fp32 columns) while it streams
// its share of the pixel tiles (split-K across CTAs; the {0,1} group gets twice as many CTAs), then red.global-adds it.
// Cout == 64: M = 128 would waste half the instruction, so the two 64-row halves of A are the SAME 64 channels of dZ one
// image row apart (the dZ box is 17 rows; the second MN block starts 1024 bytes = one tile row later).  Against the
// same B rows the shifted half sees the halo one row further down, i.e. it accumulates the NEXT filter row:
// UMMA(B rows 2k)   -> lanes 64..127: filter row 0, lanes 0..63: filter row 1
// UMMA(B rows 2k+1) -> lanes 64..127: filter row 1 (again, not drained), lanes 0..63: filter row 2
// Two instructions instead of three cover all nine taps and one CTA owns the whole dW block (n1 = 0).  The shifted
// half walks image rows ty*16-1 .. ty*16+14, so the tile grid has ceil((H+1)/16) rows.
#pragma once
#include <cuda_bf16.h>

#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kWgThreads = 192;            // TMA warp, MMA warp, 4 drain warps
constexpr int kWgDzBytes = 128 * 128;      // one 64-channel dZ box
constexpr int kWgHaloBox = 180 * 128;
constexpr int kWgHaloBuf = 23 * 1024;
constexpr int kWgStageBytes = 2 * kWgDzBytes + kWgHaloBuf;   // 55 KB (XB = 128); the XB = 32 variant uses 2*16 KB + 6 KB

struct WgradParams {
  CUtensorMap tm_dz;     // (Cout, W, H, B) bf16, box (64, 8, 16, 1)
  CUtensorMap tm_x0;     // (C0, W, H, B) bf16, box (64, 10, 18, 1)
  CUtensorMap tm_x1;     // second source of the virtual concat
  float* dw;             // [Cout][9][C0+C1] fp32, accumulated
  int cb0, cb1;          // 64-channel blocks of the two X sources
  int off_x, off_y;
  int tiles_x, tiles_y, batch;
  int Cout;
  int co_blocks;         // ceil(Cout / 128)
  int split;             // XB = 32: CTAs per dW block (split-K over the pixel tiles)
  int n0, n1;            // XB = 128: CTAs per dW block for filter rows {0,1} / for filter row {2}
  int blocks;            // co_blocks * (cb0 + cb1)
  int stages;
  int flags;             // bit 0 (experiments only): skip the drain
  int tw16;              // XB = 128: pixel tile is 16 wide x 8 tall instead of 8 x 16 (fewer padded rows on 20x26 / 40x53)
};

// MN-major, 128B-swizzled operand descriptor: LBO = distance between 64-element MN blocks, SBO = distance between
// 8-row K atoms (both >> 4); version 1, SWIZZLE_128B.
__device__ __forceinline__ uint32_t mn_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t mn_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}

__device__ __forceinline__ void red_add_v4(float* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d)) : "memory");
}

// XB = bytes of one halo pixel record of X: 128 (64 channels, N = 64 per tap, taps split into 2 groups across CTAs)
// or 32 (the 16-channel padded network input of the first conv: N = 16 per tap, all 9 taps in one CTA).
template <int XB>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  constexpr int NT = XB / 2;                                   // accumulator columns per tap
  constexpr int HALO_BOX = 180 * XB;
  constexpr int HALO_BUF = (HALO_BOX + 1023) / 1024 * 1024;
  constexpr int STAGE = 2 * kWgDzBytes + HALO_BUF;
  constexpr uint32_t XLAYOUT = (XB == 128) ? 2u : 6u;           // SWIZZLE_128B / SWIZZLE_32B
  constexpr int TCOLS = (XB == 128) ? 512 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t s_stage = smem_base;
  const uint32_t s_bar = s_stage + stages * STAGE;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * stages, bar_done = bar_empty + 8 * stages;
  const uint32_t s_tmem_slot = bar_done + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_dz);
    tma_prefetch_desc(&p.tm_x0);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TCOLS>(s_tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // block decode.  XB = 128: the blocks*n0 two-row work items come first (they are the expensive ones: the block
  // scheduler hands later CTAs to whichever SM frees up, so longest-first keeps the tail short), then the blocks*n1
  // one-row items;  XB = 32: blockIdx.x = block * split + s, all 9 taps.
  const int cbt = p.cb0 + p.cb1;
  int bid = blockIdx.x;
  int s, tg = 0, nsplit = p.split;
  if (XB == 128) {
    if (bid < p.blocks * p.n0) { tg = 0; nsplit = p.n0; }
    else { tg = 1; nsplit = p.n1; bid -= p.blocks * p.n0; }
    s = bid % nsplit;
    bid /= nsplit;
  } else {
    s = bid % p.split;
    bid /= p.split;
  }
  const int ci_blk = bid % cbt;
  const int co_blk = bid / cbt;
  const int dy0 = tg ? 2 : 0, ndy = tg ? 1 : 2;                 // XB = 128 only
  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;
  const bool half_m = (p.Cout - co_blk * 128) < 128;      // only 64 real co rows
  const bool shifted = half_m && XB == 128;                // ... used twice, one image row apart (see above)

  // pixel tile: 8 wide x 16 tall, or 16 x 8 (p.tw16); either way 128 pixels = 16 K atoms of 8 consecutive pixels
  const int TW = (XB == 128 && p.tw16) ? 16 : 8, TH = 128 / TW;
  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int t = s; t < m_tiles; t += nsplit) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        mbar_wait(bar_empty + 8 * st, ph ^ 1);
        const uint32_t sa = s_stage + st * STAGE, fb = bar_full + 8 * st;
        mbar_arrive_expect_tx(fb, (shifted ? kWgDzBytes + TW * 128 : half_m ? kWgDzBytes : 2 * kWgDzBytes) + HALO_BOX);
        if (shifted) {
          tma_load_4d(sa, &p.tm_dz, fb, co_blk * 128, tx * TW, ty * TH - 1, b);       // (TH + 1)-row box (host)
        } else {
          tma_load_4d(sa, &p.tm_dz, fb, co_blk * 128, tx * TW, ty * TH, b);
          if (!half_m) tma_load_4d(sa + kWgDzBytes, &p.tm_dz, fb, co_blk * 128 + 64, tx * TW, ty * TH, b);
        }
        if (ci_blk < p.cb0)
          tma_load_4d(sa + 2 * kWgDzBytes, &p.tm_x0, fb, ci_blk * NT, tx * TW - 1, ty * TH - 1, b);
        else
          tma_load_4d(sa + 2 * kWgDzBytes, &p.tm_x1, fb, (ci_blk - p.cb0) * NT, tx * TW - 1 - p.off_x, ty * TH - 1 - p.off_y, b);
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: fp32 accum, bf16 A/B, A and B MN-major (bits 15, 16), N = NT, M = 128
    constexpr uint32_t b_hi = ((10u * XB) >> 4) | (1u << 14) | (XLAYOUT << 29);   // K atoms = image rows, 10 halo pixels apart
    constexpr uint32_t idesc48 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((48u >> 3) << 17) | ((128u >> 4) << 24);
    // N = 192: the three dx taps of a filter row in one instruction (XB = 128)
    constexpr uint32_t idesc3 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((192u >> 3) << 17) | ((128u >> 4) << 24);
    int st = 0; uint32_t ph = 0;
    bool first = true;
    for (int t = s; t < m_tiles; t += nsplit) {
      mbar_wait(bar_full + 8 * st, ph);
      tc_fence_after();
      const uint32_t sa = s_stage + st * STAGE;
      // Cout == 64: second half = dZ one image row later (XB = 128) / the same 64 rows again (XB = 32)
      const uint32_t a_lbo = shifted ? (uint32_t)(TW * 128) : half_m ? 0u : (uint32_t)kWgDzBytes;
      if (elect_one()) {
        if (XB == 128) {
          // UMMA k covers pixels 16k..16k+15 of the tile = two 8-pixel K atoms: image rows 2k, 2k+1 of an 8-wide tile
          // (atoms one halo row = 10 pixels apart) or the two halves of row k of a 16-wide tile (atoms 8 pixels apart)
          const uint32_t b_hi_rt = p.tw16 ? mn_desc_hi(1024) : b_hi;
          const int row_mul = p.tw16 ? 1 : 2, row_pitch = p.tw16 ? 18 : 10;
          for (int dyi = 0; dyi < ndy; ++dyi) {
            const uint32_t d_tmem = tmem_base + dyi * 192;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t a_lo = mn_desc_lo(sa + k * 2048, a_lbo);
              // three MN blocks (dx = 0, 1, 2) one halo pixel = 128 bytes apart
              const uint32_t b_lo = mn_desc_lo(sa + 2 * kWgDzBytes + ((row_mul * k + dy0 + dyi) * row_pitch) * 128, 128);
              umma_bf16_lohi(d_tmem, a_lo, mn_desc_hi(1024), b_lo, b_hi_rt, idesc3, (first && k == 0) ? 0u : 1u);
            }
          }
        } else {
          // XB = 32: N = 3 x 16, the three dx taps are MN blocks one 32-byte halo pixel apart
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t d_tmem = tmem_base + dy * 3 * NT;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t a_lo = mn_desc_lo(sa + k * 2048, a_lbo);
              const uint32_t b_lo = mn_desc_lo(sa + 2 * kWgDzBytes + ((2 * k + dy) * 10) * XB, XB);
              umma_bf16_lohi(d_tmem, a_lo, mn_desc_hi(1024), b_lo, b_hi, idesc48, (first && k == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(bar_empty + 8 * st);
      }
      __syncwarp();
      first = false;
      if (++st == stages) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
  } else {
    // drain: TMEM -> red.global.add.f32 into dW[co][tap][ci]
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // co within the block
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int co = co_blk * 128 + row;
    const int ctot = cbt * NT;
    const bool live = (s < m_tiles) && (shifted || (co < p.Cout && !(half_m && row >= 64))) && !(p.flags & 1);
    if (XB == 128) {
      for (int tp = 0; tp < 3 * ndy; ++tp) {         // accumulator column block tp*64 <-> tap (dy0 + tp/3, tp%3)
        // Cout == 64: lanes 0..63 hold filter rows {1, 2}, lanes 64..127 rows {0, 1}; row 1 is drained once
        const int tap = shifted ? tp + (q < 2 ? 3 : 0) : dy0 * 3 + tp;
        if (shifted && q >= 2 && tp >= 3) continue;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + tp * 64 + c0 + ((uint32_t)(q * 32) << 16), v);
          tmem_ld_wait();
          if (live) {
            float* dst = p.dw + ((size_t)(shifted ? (co & 63) : co) * 9 + tap) * ctot + ci_blk * 64 + c0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) red_add_v4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);   // 16-byte aligned
          }
        }
      }
    } else {
      // 9 taps x 16 columns = 144 accumulator columns: 32-column loads cover two taps each (the last one half)
#pragma unroll 1
      for (int c0 = 0; c0 < 160; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + c0 + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c0 + i;
            if (col < 144) atomicAdd(p.dw + ((size_t)co * 9 + (col >> 4)) * 16 + (col & 15), __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TCOLS>(tmem_base);
}

}  // namespace gsd

// ------------------------------------------------------------------------------------------------
// Transposed-conv (k=2, s=2) weight gradient, same GEMM-over-pixels scheme without the halo:
//   dWt[ci][co][gy][gx] += sum_{b,y,x} IN[b,y,x,ci] * dU[b, 2y+gy, 2x+gx, co]
// A = IN tile (M = 128 ci = two 64-channel boxes), B_g = the g-th stride-2 view of dU (N = 64 co), 4 accumulators.
namespace gsd {

constexpr int kWpStageBytes = 2 * kWgDzBytes + 4 * kWgDzBytes;   // 96 KB

struct WgradPwParams {
  CUtensorMap tm_in;      // (Cin, W, H, B), box (64, 8, 16, 1)
  CUtensorMap tm_du[4];   // stride-2 views of dU: (Cout, W, H, B), box (64, 8, 16, 1)
  float* dw;              // (Cin, Cout, 2, 2) fp32, accumulated
  int tiles_x, tiles_y, batch;
  int Cin, Cout;
  int split;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_pw_kernel(const __grid_constant__ WgradPwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int stages = 2;
  const uint32_t s_stage = smem_base;
  const uint32_t s_bar = s_stage + stages * kWpStageBytes;
  const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * stages, bar_done = bar_empty + 8 * stages;
  const uint32_t s_tmem_slot = bar_done + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(s_tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // blockIdx.x = (ci_blk * co_blocks + co_blk) * split + s
  const int co_blocks = p.Cout / 64;
  int bid = blockIdx.x;
  const int s = bid % p.split; bid /= p.split;
  const int co_blk = bid % co_blocks;
  const int ci_blk = bid / co_blocks;
  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;

  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int t = s; t < m_tiles; t += p.split) {
        const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, b = t / (p.tiles_x * p.tiles_y);
        mbar_wait(bar_empty + 8 * st, ph ^ 1);
        const uint32_t sa = s_stage + st * kWpStageBytes, fb = bar_full + 8 * st;
        mbar_arrive_expect_tx(fb, kWpStageBytes);
        tma_load_4d(sa, &p.tm_in, fb, ci_blk * 128, tx * 8, ty * 16, b);
        tma_load_4d(sa + kWgDzBytes, &p.tm_in, fb, ci_blk * 128 + 64, tx * 8, ty * 16, b);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_4d(sa + (2 + g) * kWgDzBytes, &p.tm_du[g], fb, co_blk * 64, tx * 8, ty * 16, b);
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ONE N = 256 UMMA per K step: the four stride-2 views of dU are four 64-column MN blocks exactly one 16 KB tile
    // apart (= the descriptor's leading-byte-offset), so A is read once for all four (dy,dx) groups
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    int st = 0; uint32_t ph = 0;
    bool first = true;
    for (int t = s; t < m_tiles; t += p.split) {
      mbar_wait(bar_full + 8 * st, ph);
      tc_fence_after();
      const uint32_t sa = s_stage + st * kWpStageBytes;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t a_lo = mn_desc_lo(sa + k * 2048, kWgDzBytes);
          const uint32_t b_lo = mn_desc_lo(sa + 2 * kWgDzBytes + k * 2048, kWgDzBytes);
          umma_bf16_lohi(tmem_base, a_lo, mn_desc_hi(1024), b_lo, mn_desc_hi(1024), idesc, (first && k == 0) ? 0u : 1u);
        }
        umma_commit(bar_empty + 8 * st);
      }
      __syncwarp();
      first = false;
      if (++st == stages) { st = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int ci = ci_blk * 128 + q * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const bool live = s < m_tiles && ci < p.Cin;
    for (int g = 0; g < 4; ++g) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + g * 64 + c0 + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        if (live) {
          float* dst = p.dw + ((size_t)ci * p.Cout + co_blk * 64 + c0) * 4 + g;
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + 4 * i, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem_base);
}

}  // namespace gsd
