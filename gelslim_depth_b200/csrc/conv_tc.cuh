// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a), NHWC bf16.
//
//   D[pixel, n] = sum_{tap, c} X[pixel + tap, c] * Wt[n, tap, c]          (fp32 accumulate in TMEM)
//
// GEMM view: M = pixels (one CTA tile = a TH x TW spatial patch = 128 pixels = the 128 TMEM lanes),
// N = output channels (BN per tile), K = taps x input channels, walked in K-blocks of one tap x
// (BKB/2) channels.  The A operand of a K-block is ONE 4-D TMA box (channels, TW, TH, 1) of the NHWC
// activation tensor, started at the tap-shifted coordinate: TMA's out-of-bounds zero fill supplies the
// conv padding, the bottom/right F.pad of `Up` (unet.py:43-47) and ragged tile edges for free, and the
// box lands in shared memory already in the K-major swizzled layout tcgen05.mma reads.  A second
// activation source makes torch.cat([skip, up], 1) (unet.py:48) virtual: its channels are just more
// K-blocks.  Nothing is im2col-materialised.
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles):
//   warp 0    : TMA producer (one elected lane)
//   warp 1    : tcgen05.mma issuer (one elected lane); accumulators double-buffered in TMEM so the
//               epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5 : epilogue: tcgen05.ld -> per-channel scale/shift (folded BatchNorm or bias) -> ReLU ->
//               bf16 -> swizzled smem staging -> TMA store (hardware clips ragged edges);
//               optionally a fused 2x2 max-pool of the staged tile -> second TMA store.
#pragma once
#include <cuda_bf16.h>

#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kConvThreads = 192;
constexpr int kMaxTaps = 9;

struct ConvParams {
  CUtensorMap tm_src0;    // (C0, W, H, B) bf16
  CUtensorMap tm_src1;    // (C1, W1, H1, B) bf16 -- second half of the virtual concat (unused if kb1 == 0)
  CUtensorMap tm_w;       // (Ktot, Ntot) bf16, K-major weights
  CUtensorMap tm_out[4];  // (Cout, W, H, B) bf16; 4 strided views for the transposed-conv scatter
  CUtensorMap tm_pool;    // (Cout, W/2, H/2, B) bf16
  const float* scale;     // [Ntot]
  const float* shift;     // [Ntot]
  int kb0, kb1;           // channel blocks per tap of source 0 / 1
  int ntaps;
  int off_x, off_y;       // where source 1's (0,0) sits in the output frame (F.pad left/top)
  int tiles_x, tiles_y, batch;
  int th, tw;             // th * tw == 128
  int n_tiles;            // Ntot / BN
  int cout_per_group;     // Cout of ONE output view; Ntot = groups * cout_per_group
  int relu;
  int do_pool;
  int8_t tap_dy[kMaxTaps];
  int8_t tap_dx[kMaxTaps];
};

template <int BN, int BKB>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * BKB;
  static constexpr int B_BYTES = BN * BKB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int GROUPS = BN / 64;                   // 64-channel (128-byte) output groups
  static constexpr int PASS_GROUPS = GROUPS > 2 ? 2 : GROUPS;   // groups staged per epilogue pass
  static constexpr int PASSES = GROUPS / PASS_GROUPS;
  static constexpr int PASS_N = PASS_GROUPS * 64;
  static constexpr int OUT_BYTES = PASS_GROUPS * 128 * 128;     // staged bf16 tile (one pass)
  static constexpr int POOL_BYTES = PASS_GROUPS * 32 * 128;     // staged pooled tile (one pass)
  static constexpr int AUX_BYTES = 2 * BN * 4 + 256;       // scale/shift + barriers + tmem slot
  static constexpr int BUDGET = 227 * 1024 - 1024;         // minus manual 1024-byte alignment slack
  static constexpr int STAGES_RAW = (BUDGET - OUT_BYTES - POOL_BYTES - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + POOL_BYTES + AUX_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be 64, 128 or 256");
};

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a 128B-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int BN, int BKB>
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, BKB>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int KSTEPS = BKB / 32;   // tcgen05.mma instructions (K = 16 bf16 = 32 bytes) per K-block

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t s_stage = smem_base;
  const uint32_t s_out = s_stage + STAGES * Cfg::STAGE_BYTES;
  const uint32_t s_pool = s_out + Cfg::OUT_BYTES;
  const uint32_t s_aux = s_pool + Cfg::POOL_BYTES;
  float* g_scale = reinterpret_cast<float*>(smem_gen + (s_aux - smem_base));
  float* g_shift = g_scale + BN;
  const uint32_t s_bar = s_aux + 2 * BN * 4;
  // barrier slots (8 bytes each): full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
  const uint32_t bar_full = s_bar;
  const uint32_t bar_empty = s_bar + 8 * STAGES;
  const uint32_t bar_acc_full = s_bar + 16 * STAGES;
  const uint32_t bar_acc_empty = bar_acc_full + 16;
  const uint32_t s_tmem_slot = bar_acc_empty + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_src0);
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_out[0]);
    if (p.kb1) tma_prefetch_desc(&p.tm_src1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(s_tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;
  const int total_tiles = m_tiles * p.n_tiles;
  const int kb_per_tap = p.kb0 + p.kb1;
  const int num_kb = p.ntaps * kb_per_tap;
  constexpr int KELEMS = BKB / 2;   // channels per K-block

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int tx = mt % p.tiles_x;
        mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int b = mt / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th;
        int kidx = 0;
        for (int t = 0; t < p.ntaps; ++t) {
          const int xs = x0 + p.tap_dx[t], ys = y0 + p.tap_dy[t];
          for (int cb = 0; cb < kb_per_tap; ++cb, ++kidx) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t sa = s_stage + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + Cfg::A_BYTES;
            const uint32_t fb = bar_full + 8 * stage;
            mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES);
            if (cb < p.kb0)
              tma_load_4d(sa, &p.tm_src0, fb, cb * KELEMS, xs, ys, b);
            else
              tma_load_4d(sa, &p.tm_src1, fb, (cb - p.kb0) * KELEMS, xs - p.off_x, ys - p.off_y, b);
            tma_load_2d(sb, &p.tm_w, fb, kidx * KELEMS, nt * BN);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (warp converged; one elected lane issues)
    constexpr uint32_t idesc = make_idesc_bf16_m128(BN);
    constexpr uint32_t desc_hi = ((8u * BKB) >> 4) | (1u << 14) | ((BKB == 128 ? 2u : 6u) << 29);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        const uint32_t sa = s_stage + stage * Cfg::STAGE_BYTES;
        const uint32_t a_lo = ((sa & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo = (((sa + Cfg::A_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle span: +2 in the >>4 address field
            umma_bf16_lohi(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (k != 0) ? 1u : (uint32_t)(kb != 0));
          }
          umma_commit(bar_empty + 8 * stage);   // frees the smem stage once these MMAs retire
          if (kb == num_kb - 1) umma_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5, 128 threads)
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;        // accumulator row == pixel index inside the tile
    const int et = threadIdx.x - 64;      // 0..127
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int b = mt / p.tiles_y;
      const int x0 = tx * p.tw, y0 = ty * p.th;
      const int n0 = nt * BN;

      const int grp_idx = n0 / p.cout_per_group;            // which output view (transposed-conv scatter)
      const int ch0 = n0 - grp_idx * p.cout_per_group;
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int pass = 0; pass < Cfg::PASSES; ++pass) {
        // staging buffers are free once the previous TMA stores have finished reading them
        if (et == 0) tma_store_wait_read<0>();
        if (pass == 0) {
          for (int i = et; i < BN; i += 128) {
            g_scale[i] = __ldg(p.scale + n0 + i);
            g_shift[i] = __ldg(p.shift + n0 + i);
          }
        }
        named_bar_sync(1, 128);
        if (pass == 0) {
          mbar_wait(bar_acc_full + 8 * acc, acc_phase);
          tc_fence_after();
        }
#pragma unroll 1
        for (int cc = 0; cc < Cfg::PASS_N; cc += 32) {
          const int c0 = pass * Cfg::PASS_N + cc;
          uint32_t v[32];
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(v[2 * j]) * g_scale[c0 + 2 * j] + g_shift[c0 + 2 * j];
            float c = __uint_as_float(v[2 * j + 1]) * g_scale[c0 + 2 * j + 1] + g_shift[c0 + 2 * j + 1];
            if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
            __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          const uint32_t grp = s_out + (cc >> 6) * (128 * 128);
          const int chunk0 = (cc & 63) >> 3;   // 0 or 4
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = grp + sw128_off(row, chunk0 + j);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * j]),
                         "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                         : "memory");
          }
        }
        if (pass == Cfg::PASSES - 1) {
          // accumulator fully read -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        }

        if (p.do_pool) {
          named_bar_sync(1, 128);   // whole staged tile visible to all epilogue threads
          const int pw = p.tw >> 1;
          for (int item = et; item < Cfg::PASS_GROUPS * 32 * 8; item += 128) {
            const int chunk = item & 7;
            const int prow = (item >> 3) & 31;
            const int g = item >> 8;
            const int py = prow / pw, px = prow - py * pw;
            const int r00 = (2 * py) * p.tw + 2 * px;
            const uint32_t gb = s_out + g * (128 * 128);
            uint4 a, c, d, e;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(gb + sw128_off(r00, chunk)));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "r"(gb + sw128_off(r00 + 1, chunk)));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "r"(gb + sw128_off(r00 + p.tw, chunk)));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "r"(gb + sw128_off(r00 + p.tw + 1, chunk)));
            auto mx = [](uint32_t u0, uint32_t u1, uint32_t u2, uint32_t u3) {
              __nv_bfloat162 m = __hmax2(__hmax2(*reinterpret_cast<__nv_bfloat162*>(&u0), *reinterpret_cast<__nv_bfloat162*>(&u1)),
                                         __hmax2(*reinterpret_cast<__nv_bfloat162*>(&u2), *reinterpret_cast<__nv_bfloat162*>(&u3)));
              return *reinterpret_cast<uint32_t*>(&m);
            };
            const uint32_t o0 = mx(a.x, c.x, d.x, e.x), o1 = mx(a.y, c.y, d.y, e.y);
            const uint32_t o2 = mx(a.z, c.z, d.z, e.z), o3 = mx(a.w, c.w, d.w, e.w);
            const uint32_t addr = s_pool + g * (32 * 128) + sw128_off(prow, chunk);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
          }
        }
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
        named_bar_sync(1, 128);
        if (et == 0) {
#pragma unroll
          for (int g = 0; g < Cfg::PASS_GROUPS; ++g) {
            const int ch = ch0 + pass * Cfg::PASS_N + g * 64;
            tma_store_4d(&p.tm_out[grp_idx], s_out + g * (128 * 128), ch, x0, y0, b);
            if (p.do_pool) tma_store_4d(&p.tm_pool, s_pool + g * (32 * 128), ch, x0 >> 1, y0 >> 1, b);
          }
          tma_store_commit();
        }
      }
    }
    if (et == 0) tma_store_wait_all<0>();   // global writes complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace gsd
