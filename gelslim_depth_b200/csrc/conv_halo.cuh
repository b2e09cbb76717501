// 3x3 convolution as implicit GEMM on tcgen05 with the input HALO kept in shared memory.
//
// Why: conv_tc.cuh re-loads the 128-pixel A tile once per filter tap (9x) and the weight tile once per
// output tile; measured on B200 every layer then sits on the L2->SM fabric limit (~50 B/clk/SM), not on
// the tensor pipe.  Here one TMA box per 64-channel block brings the (16+2) x (8+2) pixel halo of a
// 16 x 8 output tile into smem ONCE; the A operand of filter tap (dy,dx) is the SAME smem buffer read
// through a UMMA descriptor whose start address is shifted by (dy*10+dx) rows: the 16 eight-row core
// groups (one per image row of the tile) are 10 rows = 1280 bytes apart (stride-byte-offset), and the
// descriptor's base-offset field carries the 128B-swizzle phase of the unaligned start row.  A-operand
// traffic from L2 drops 6.4x.  Weights are either RESIDENT in smem for the whole persistent CTA (64/128
// output channels with small K) or streamed through a ring and shared by MT=2 output tiles (M = 256
// per CTA), which halves their traffic.
//
// Epilogue: tcgen05.ld -> scale/shift (folded BatchNorm) -> ReLU -> bf16 -> 16-byte global stores
// straight from registers (each thread owns one pixel = 128 contiguous bytes per 64 channels); the
// fused 2x2 max-pool is two warp shuffles (the 4 pixels of a window live in lanes l, l^1, l^8, l^9).
#pragma once
#include <cuda_bf16.h>

#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kHaloThreads = 192;
constexpr int kHaloRows = 18 * 10;                 // halo pixels per tile
constexpr int kHaloBoxBytes = kHaloRows * 128;     // 23040: bytes one TMA box delivers
constexpr int kHaloBufBytes = 23 * 1024;           // buffer pitch (1024-aligned)

struct HaloParams {
  CUtensorMap tm_src0;   // (C0, W, H, B) bf16, box (64, 10, 18, 1)
  CUtensorMap tm_src1;   // second source of the virtual concat
  CUtensorMap tm_w;      // (9*(C0+C1), Cout) bf16, box (64, BN)
  const float* scale;    // [Cout]
  const float* shift;    // [Cout]
  __nv_bfloat16* out;    // (B, H, W, Cout)
  __nv_bfloat16* pooled; // (B, H/2, W/2, Cout) or null
  int cb0, cb1;          // 64-channel blocks of source 0 / 1
  int off_x, off_y;      // F.pad left/top of source 1
  int tiles_x, tiles_y, batch;
  int H, W, Cout;
  int n_tiles;           // Cout / BN
  int relu;
  int na, nb;            // ring depths (A halo buffers, B weight stages)
  int base_off_mode;     // 1: descriptor base-offset = (start >> 7) & 7 (PTX ISA); 0: leave it 0 (experiment)
};

// high words of the two smem descriptors (constant): version 1 @bit 46, SW128 @bits 61..63, SBO >> 4 @bits 32..45
constexpr uint32_t kHaloDescHi = (1280u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kKmajor128DescHi = (1024u >> 4) | (1u << 14) | (2u << 29);

// smem descriptor for the shifted-halo A operand
__device__ __forceinline__ uint64_t make_halo_desc(uint32_t saddr, int base_off_mode) {
  const uint64_t bo = base_off_mode ? (uint64_t)((saddr >> 7) & 7u) : 0ull;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1280 >> 4) << 32) | (1ull << 46) | (bo << 49) |
         (2ull << 61);
}

template <int BN, int MT, bool WRES>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int B_BYTES = BN * 128;
  constexpr int TMEM_COLS = (2 * MT * BN <= 128) ? 128 : (2 * MT * BN <= 256) ? 256 : 512;
  static_assert(2 * MT * BN <= 512, "accumulators exceed TMEM");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int cbt = p.cb0 + p.cb1;
  const int nkb = 9 * cbt;
  const int na = p.na, nb = WRES ? 0 : p.nb;
  const uint32_t s_a = smem_base;
  const uint32_t s_b = s_a + na * kHaloBufBytes;                       // B ring, or the resident weights
  const uint32_t s_aux = s_b + (WRES ? nkb : nb) * B_BYTES;
  float* g_scale = reinterpret_cast<float*>(smem_gen + (s_aux - smem_base));   // [Cout] (<= 1024)
  float* g_shift = g_scale + p.Cout;
  const uint32_t s_bar = s_aux + 2 * p.Cout * 4;
  const uint32_t bar_fullA = s_bar;                    // [na]
  const uint32_t bar_emptyA = bar_fullA + 8 * na;      // [na]
  const uint32_t bar_fullB = bar_emptyA + 8 * na;      // [nb] (or [1] = resident weights landed)
  const uint32_t bar_emptyB = bar_fullB + 8 * (WRES ? 1 : nb);
  const uint32_t bar_acc_full = bar_emptyB + 8 * (WRES ? 1 : nb);   // [2]
  const uint32_t bar_acc_empty = bar_acc_full + 16;                  // [2]
  const uint32_t s_tmem_slot = bar_acc_empty + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_src0);
    tma_prefetch_desc(&p.tm_w);
    if (p.cb1) tma_prefetch_desc(&p.tm_src1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) { mbar_init(bar_fullA + 8 * i, 1); mbar_init(bar_emptyA + 8 * i, 1); }
    for (int i = 0; i < (WRES ? 1 : nb); ++i) { mbar_init(bar_fullB + 8 * i, 1); mbar_init(bar_emptyB + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(s_tmem_slot);
  for (int i = threadIdx.x; i < p.Cout; i += kHaloThreads) {
    g_scale[i] = __ldg(p.scale + i);
    g_shift[i] = __ldg(p.shift + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;
  const int m_groups = (m_tiles + MT - 1) / MT;
  const int total_items = m_groups * p.n_tiles;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      if (WRES) {
        mbar_arrive_expect_tx(bar_fullB, (uint32_t)(nkb * B_BYTES));
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(s_b + kb * B_BYTES, &p.tm_w, bar_fullB, kb * 64, 0);
      }
      int ia = 0, ib = 0;
      uint32_t pa = 0, pb = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int nt = item % p.n_tiles;
        const int mg = item / p.n_tiles;
        for (int cb = 0; cb < cbt; ++cb) {
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            int mt = mg * MT + j;
            if (mt >= m_tiles) mt = m_tiles - 1;      // dummy duplicate keeps the pipeline protocol uniform
            const int tx = mt % p.tiles_x;
            const int ty = (mt / p.tiles_x) % p.tiles_y;
            const int b = mt / (p.tiles_x * p.tiles_y);
            const int xs = tx * 8 - 1, ys = ty * 16 - 1;
            mbar_wait(bar_emptyA + 8 * ia, pa ^ 1);
            mbar_arrive_expect_tx(bar_fullA + 8 * ia, kHaloBoxBytes);
            if (cb < p.cb0)
              tma_load_4d(s_a + ia * kHaloBufBytes, &p.tm_src0, bar_fullA + 8 * ia, cb * 64, xs, ys, b);
            else
              tma_load_4d(s_a + ia * kHaloBufBytes, &p.tm_src1, bar_fullA + 8 * ia, (cb - p.cb0) * 64, xs - p.off_x,
                          ys - p.off_y, b);
            if (++ia == na) { ia = 0; pa ^= 1; }
          }
          if (!WRES) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(bar_emptyB + 8 * ib, pb ^ 1);
              mbar_arrive_expect_tx(bar_fullB + 8 * ib, B_BYTES);
              tma_load_2d(s_b + ib * B_BYTES, &p.tm_w, bar_fullB + 8 * ib, (tap * cbt + cb) * 64, nt * BN);
              if (++ib == nb) { ib = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (whole warp converged; one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16_m128(BN);
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    if (WRES) { mbar_wait(bar_fullB, 0); tc_fence_after(); }
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(bar_acc_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int cb = 0; cb < cbt; ++cb) {
        uint32_t a_lo[MT];     // low word of the halo descriptor at tap (0,0), k = 0
        int a_slot[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          mbar_wait(bar_fullA + 8 * ia, pa);
          a_lo[j] = ((s_a + ia * kHaloBufBytes) & 0x3FFFFu) >> 4;
          a_slot[j] = ia;
          if (++ia == na) { ia = 0; pa ^= 1; }
        }
        tc_fence_after();
#pragma unroll 1
        for (int ty3 = 0; ty3 < 3; ++ty3) {
#pragma unroll
          for (int tx3 = 0; tx3 < 3; ++tx3) {
            const int tap = ty3 * 3 + tx3;
            uint32_t b_lo;
            if (WRES) {
              b_lo = ((s_b + (tap * cbt + cb) * B_BYTES) & 0x3FFFFu) >> 4;
            } else {
              mbar_wait(bar_fullB + 8 * ib, pb);
              tc_fence_after();
              b_lo = ((s_b + ib * B_BYTES) & 0x3FFFFu) >> 4;
            }
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < MT; ++j) {
                const uint32_t d_tmem = tmem_base + (buf * MT + j) * BN;
                const uint32_t a0 = a_lo[j] + (ty3 * 10 + tx3) * 8;     // +128 bytes per halo row (>>4)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_lohi(d_tmem, (a0 + 2 * k) | (1u << 16), kHaloDescHi, (b_lo + 2 * k) | (1u << 16), kKmajor128DescHi,
                                 idesc, (k != 0) ? 1u : (uint32_t)((cb | tap) != 0));
              }
              if (!WRES) umma_commit(bar_emptyB + 8 * ib);
            }
            __syncwarp();
            if (!WRES) { if (++ib == nb) { ib = 0; pb ^= 1; } }
          }
        }
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < MT; ++j) umma_commit(bar_emptyA + 8 * a_slot[j]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar_acc_full + 8 * buf);
      __syncwarp();
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;
    const int ly = 4 * q + (lane >> 3), lx = lane & 7;
    const int Hp = p.H >> 1, Wp = p.W >> 1;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const int nt = item % p.n_tiles;
      const int mg = item / p.n_tiles;
      mbar_wait(bar_acc_full + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int mt = mg * MT + j;
        if (mt >= m_tiles) break;           // warp-uniform
        const int tx = mt % p.tiles_x;
        const int ty = (mt / p.tiles_x) % p.tiles_y;
        const int b = mt / (p.tiles_x * p.tiles_y);
        const int y = ty * 16 + ly, x = tx * 8 + lx;
        const bool valid = (y < p.H) && (x < p.W);
        const bool pvalid = p.pooled && !(lane & 9) && ((y >> 1) < Hp) && ((x >> 1) < Wp);
        __nv_bfloat16* orow = p.out + (((size_t)b * p.H + y) * p.W + x) * p.Cout + nt * BN;
        __nv_bfloat16* prow = p.pooled ? p.pooled + (((size_t)b * Hp + (y >> 1)) * Wp + (x >> 1)) * p.Cout + nt * BN : nullptr;
        const uint32_t t_row = tmem_base + (buf * MT + j) * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
          const float* sc = g_scale + nt * BN + c0;
          const float* sh = g_shift + nt * BN + c0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(v[2 * i]) * sc[2 * i] + sh[2 * i];
            float c = __uint_as_float(v[2 * i + 1]) * sc[2 * i + 1] + sh[2 * i + 1];
            if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (valid) {
            uint4* o = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
          if (p.pooled) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              uint32_t o1 = __shfl_xor_sync(0xffffffffu, pk[i], 1);
              __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pk[i]), *reinterpret_cast<__nv_bfloat162*>(&o1));
              uint32_t mu = *reinterpret_cast<uint32_t*>(&m);
              uint32_t o8 = __shfl_xor_sync(0xffffffffu, mu, 8);
              m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o8));
              pk[i] = *reinterpret_cast<uint32_t*>(&m);
            }
            if (pvalid) {
              uint4* o = reinterpret_cast<uint4*>(prow + c0);
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace gsd
