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
// Warp roles (320 threads, 1 CTA / SM, persistent over tiles):
//   warp 0    : TMA producer (one elected lane)
//   warp 1    : tcgen05.mma issuer (one elected lane); accumulators double-buffered in TMEM so the
//               epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-9 : epilogue (epilogue.cuh): tcgen05.ld -> per-channel scale/shift (folded BatchNorm or bias) ->
//               ReLU -> bf16 -> 16-byte global stores from registers; optional fused 2x2 max-pool by warp
//               shuffles; the transposed-conv variant scatters to (2y+dy, 2x+dx).
#pragma once
#include <cuda_bf16.h>

#include "epilogue.cuh"
#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kConvThreads = 64 + 32 * kEpiWarps;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int kMaxTaps = 9;

struct ConvParams {
  CUtensorMap tm_src0;    // (C0, W, H, B) bf16
  CUtensorMap tm_src1;    // (C1, W1, H1, B) bf16 -- second half of the virtual concat (unused if kb1 == 0)
  CUtensorMap tm_w;       // (Ktot, Ntot) bf16, K-major weights
  __nv_bfloat16* out;     // (B, H, W, Cout), or (B, 2H, 2W, Cout) for the transposed-conv scatter (groups == 4)
  __nv_bfloat16* pooled;  // (B, H/2, W/2, Cout) or null
  int H, W, groups, ntot;
  float* stats;           // global [2][ntot] batch statistics of the raw conv output, or null
  FastDiv fd_ntiles, fd_tx, fd_ty, fd_cpg;   // fd_cpg: division by cout_per_group
  int src5;               // 1: tm_src0 is the 5-D space-to-depth view (2C, W, 2, H, B) of a (B,2H,2W,C) tensor and the
                          //    "tap" index is its gy coordinate (transposed-conv input gradient)
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

// CTA2: CTA pair (cta_group::2, see conv_halo.cuh): each CTA stages BN / 2 weight rows per K block
template <int BN, int BKB, bool CTA2 = false>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * BKB;
  static constexpr int BN_CTA = CTA2 ? BN / 2 : BN;
  static constexpr int B_BYTES = BN_CTA * BKB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_NTOT = 2048;                    // scale/shift of the whole layer live in smem
  static constexpr int AUX_BYTES = 4 * MAX_NTOT * 4 + kEpiWarps * kEpiStageBytesPerWarp + 256;   // scale/shift/stats + patches + barriers
  static constexpr int BUDGET = 227 * 1024 - 1024;         // minus manual 1024-byte alignment slack
  static constexpr int STAGES_RAW = (BUDGET - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(BN % 64 == 0 && BN <= 256, "BN must be 64, 128 or 256");
};

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a 128B-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int BN, int BKB, bool CTA2 = false>
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, BKB, CTA2>;
  constexpr uint32_t NCTA = CTA2 ? 2 : 1;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int KSTEPS = BKB / 32;   // tcgen05.mma instructions (K = 16 bf16 = 32 bytes) per K-block

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t s_stage = smem_base;
  const uint32_t s_aux = s_stage + STAGES * Cfg::STAGE_BYTES;
  float* g_scale = reinterpret_cast<float*>(smem_gen + (s_aux - smem_base));
  float* g_shift = g_scale + Cfg::MAX_NTOT;
  float* g_stats = g_shift + Cfg::MAX_NTOT;                 // [2][ntot]
  const uint32_t s_epi = s_aux + 4 * Cfg::MAX_NTOT * 4;
  const uint32_t s_bar = s_epi + kEpiWarps * kEpiStageBytesPerWarp;
  // barrier slots (8 bytes each): full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
  const uint32_t bar_full = s_bar;
  const uint32_t bar_empty = s_bar + 8 * STAGES;
  const uint32_t bar_acc_full = s_bar + 16 * STAGES;
  const uint32_t bar_acc_empty = bar_acc_full + 16;
  const uint32_t s_tmem_slot = bar_acc_empty + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_src0);
    tma_prefetch_desc(&p.tm_w);
    if (p.kb1) tma_prefetch_desc(&p.tm_src1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4 * NCTA);   // one arrive per epilogue warp of the set that owns the buffer (both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) { if (CTA2) tmem_alloc_2sm<Cfg::TMEM_COLS>(s_tmem_slot); else tmem_alloc<Cfg::TMEM_COLS>(s_tmem_slot); }
  for (int i = threadIdx.x; i < p.ntot; i += kConvThreads) {
    if (p.scale) g_scale[i] = __ldg(p.scale + i);      // null: scale == 1 / shift == 0, the epilogue skips the loads
    if (p.shift) g_shift[i] = __ldg(p.shift + i);
    g_stats[i] = 0.f;
    g_stats[p.ntot + i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_launch_dependents();

  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;
  const int total_tiles = (CTA2 ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;     // CTA2: work items are PAIRS of M tiles
  const int first_tile = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kb_per_tap = p.kb0 + p.kb1;
  const int num_kb = p.ntaps * kb_per_tap;
  constexpr int KELEMS = BKB / 2;   // channels per K-block

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();
      for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
        uint32_t mt, nt, trow, tx, b, ty;
        fdivmod((uint32_t)tile, p.fd_ntiles, mt, nt);
        if (CTA2) { mt = 2 * mt + rank; if ((int)mt >= m_tiles) mt = m_tiles - 1; }   // odd tail: the peer re-loads the last tile
        fdivmod(mt, p.fd_tx, trow, tx);
        fdivmod(trow, p.fd_ty, b, ty);
        const int x0 = tx * p.tw, y0 = ty * p.th;
        // K order = (channel block, tap), the same as the halo-resident kernel: a layer's result does not depend on
        // which of the two kernels the batch size selected (bit-identical accumulation order)
        for (int cb = 0; cb < kb_per_tap; ++cb) {
          for (int t = 0; t < p.ntaps; ++t) {
            const int xs = x0 + p.tap_dx[t], ys = y0 + p.tap_dy[t];
            const int kidx = t * kb_per_tap + cb;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t sa = s_stage + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + Cfg::A_BYTES;
            const uint32_t fb = bar_full + 8 * stage;
            if (leader) mbar_arrive_expect_tx(fb, Cfg::STAGE_BYTES * NCTA);   // both CTAs' boxes land on the leader's barrier
            if (CTA2) {
              if (p.src5)
                tma_load_5d_2sm(sa, &p.tm_src0, fb, cb * KELEMS, x0, t, y0, b);
              else if (cb < p.kb0)
                tma_load_4d_2sm(sa, &p.tm_src0, fb, cb * KELEMS, xs, ys, b);
              else
                tma_load_4d_2sm(sa, &p.tm_src1, fb, (cb - p.kb0) * KELEMS, xs - p.off_x, ys - p.off_y, b);
              tma_load_2d_2sm(sb, &p.tm_w, fb, kidx * KELEMS, nt * BN + (int)rank * Cfg::BN_CTA);
            } else {
              if (p.src5)
                tma_load_5d(sa, &p.tm_src0, fb, cb * KELEMS, x0, t, y0, b);
              else if (cb < p.kb0)
                tma_load_4d(sa, &p.tm_src0, fb, cb * KELEMS, xs, ys, b);
              else
                tma_load_4d(sa, &p.tm_src1, fb, (cb - p.kb0) * KELEMS, xs - p.off_x, ys - p.off_y, b);
              tma_load_2d(sb, &p.tm_w, fb, kidx * KELEMS, nt * BN);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
   if (leader) {
    // ===================================================== MMA issuer (warp converged; one elected lane issues)
    constexpr uint32_t idesc = CTA2 ? make_idesc_bf16_m256(BN) : make_idesc_bf16_m128(BN);
    constexpr uint32_t desc_hi = ((8u * BKB) >> 4) | (1u << 14) | ((BKB == 128 ? 2u : 6u) << 29);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
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
            if (CTA2) umma_bf16_lohi_2sm(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (k != 0) ? 1u : (uint32_t)(kb != 0));
            else umma_bf16_lohi(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (k != 0) ? 1u : (uint32_t)(kb != 0));
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs retire
          if (CTA2) umma_commit_2sm(bar_empty + 8 * stage); else umma_commit(bar_empty + 8 * stage);
          if (kb == num_kb - 1) { if (CTA2) umma_commit_2sm(bar_acc_full + 8 * acc); else umma_commit(bar_acc_full + 8 * acc); }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
   }
  } else {
    // ===================================================== epilogue (warps 2..9): quadrant q = warp % 4, set = (warp-2)/4
    // (the two sets of four warps drain alternate tiles = one set per TMEM accumulator buffer)
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    const int eset = (warp - 2) >> 2;
    const int ew = warp - 2;
    const int tw_shift = 31 - __clz(p.tw);          // tile width is a power of two
    const int row = q * 32 + lane;        // accumulator row == pixel index inside the tile
    const int ly = row >> tw_shift, lx = row & (p.tw - 1);
    const int Hp = p.H >> 1, Wp = p.W >> 1;
    const bool hx = lane & 1, hy = (lane & p.tw) != 0;   // pooling needs tw in {8, 16}: both window rows in one warp
    int it = 0;
    pdl_wait();
    for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
      const int acc = it & 1;
      if (acc != eset) continue;
      const uint32_t acc_phase = (it >> 1) & 1;
      uint32_t mt, nt, trow, tx, b, ty;
      fdivmod((uint32_t)tile, p.fd_ntiles, mt, nt);
      if (CTA2) mt = 2 * mt + rank;
      if (CTA2 && (int)mt >= m_tiles) {       // odd tail: this CTA's half of the pair is a duplicate -- hand the buffer back
        mbar_wait(bar_acc_full + 8 * acc, acc_phase);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_acc_empty + 8 * acc);
        continue;
      }
      fdivmod(mt, p.fd_tx, trow, tx);
      fdivmod(trow, p.fd_ty, b, ty);
      const int y = ty * p.th + ly, x = tx * p.tw + lx;
      const int n0 = nt * BN;
      EpiPixel px;
      px.store_out = true;
      px.valid = (y < p.H) && (x < p.W);
      px.s_stats = p.stats ? g_stats : nullptr;
      px.stats_ch0 = n0;
      px.stats_stride = p.ntot;
      px.pvalid = ((y >> 1) < Hp) && ((x >> 1) < Wp);
      px.hx = hx; px.hy = hy; px.ypart = p.tw;
      // &out[pixel][0] of the 4 rows this lane stores after the transpose (null: outside the image); the per-unit
      // part of the address (channel, transposed-conv (dy,dx) view) is warp-uniform and added in the loop below
      __nv_bfloat16* row_base[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = q * 32 + 8 * i + (lane >> 2);
        const int yy = ty * p.th + (rr >> tw_shift), xx = tx * p.tw + (rr & (p.tw - 1));
        if (yy >= p.H || xx >= p.W) row_base[i] = nullptr;
        else if (p.groups == 1) row_base[i] = p.out + (size_t)(((int)b * p.H + yy) * p.W + xx) * p.cout_per_group;
        else row_base[i] = p.out + (size_t)(((int)b * 2 * p.H + 2 * yy) * (2 * p.W) + 2 * xx) * p.cout_per_group;
      }
      __nv_bfloat16* const pool_base =
          p.pooled ? p.pooled + (size_t)(((int)b * Hp + (y >> 1)) * Wp + (x >> 1)) * p.cout_per_group + (hx ? 16 : 0) + (hy ? 8 : 0)
                   : nullptr;
      mbar_wait(bar_acc_full + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      float hacc[4];
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        // a 32-column unit lies inside ONE (dy,dx) group of the transposed-conv scatter (Cout % 32 == 0)
        uint32_t grp_idx, ch0;                                        // ch0: channel of column c0 inside its group
        fdivmod((uint32_t)(n0 + c0), p.fd_cpg, grp_idx, ch0);
        const int uoff = ((int)(grp_idx >> 1) * (2 * p.W) + (int)(grp_idx & 1)) * p.cout_per_group + (int)ch0 - c0;
#pragma unroll
        for (int i = 0; i < 4; ++i) px.rp[i] = row_base[i] ? row_base[i] + uoff : nullptr;
        px.prow = pool_base ? pool_base + ((int)ch0 - c0) : nullptr;
        epilogue_32cols(t_row, c0, p.scale ? g_scale + n0 : nullptr, p.shift ? g_shift + n0 : nullptr, p.relu, px, s_epi + ew * kEpiStageBytesPerWarp, lane, hacc, nullptr, 0);
      }
      // accumulator fully read -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CTA2) mbar_arrive_leader(bar_acc_empty + 8 * acc); else mbar_arrive(bar_acc_empty + 8 * acc); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();
  if (p.stats) {
    for (int i = threadIdx.x; i < 2 * p.ntot; i += kConvThreads) atomicAdd(p.stats + i, g_stats[i]);
  }
  if (warp == 2) { if (CTA2) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base); else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

}  // namespace gsd
