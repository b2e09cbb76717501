// 3x3 convolution as implicit GEMM on tcgen05 with the input HALO kept in shared memory.
//
// Why: conv_tc.cuh re-loads the 128-pixel A tile once per filter tap (9x) and the weight tile once per
// output tile; measured on B200 every layer then sits on the L2->SM fabric limit (~50 B/clk/SM), not on
// the tensor pipe.  Here one TMA box per 64-channel block brings the (16+2) x (8+2) pixel halo of a
// 16 x 8 output tile into smem ONCE; the A operand of filter tap (dy,dx) is the SAME smem buffer read
// through a UMMA descriptor whose start address is shifted by (dy*10+dx) rows: the 16 eight-row core
// groups (one per image row of the tile) are 10 rows = 1280 bytes apart (stride-byte-offset).  The start
// row is NOT 8-row aligned; measured on B200 the tensor core applies the 128B-swizzle XOR on absolute
// shared-memory address bits (as TMA does when it writes), so the descriptor base-offset stays 0.
// A-operand traffic from L2 drops 6.4x.  Weights are either RESIDENT in smem for the whole persistent CTA (64/128
// output channels with small K) or streamed through a ring and shared by MT=2 output tiles (M = 256
// per CTA), which halves their traffic.
//
// Epilogue (epilogue.cuh): tcgen05.ld -> scale/shift (folded BatchNorm, packed fp32) -> ReLU folded into the bf16
// conversion -> per-warp swizzled smem transposition -> 16-byte global stores in which four lanes cover one pixel's
// 64 bytes; the fused 2x2 max-pool is 12 warp shuffles (the 4 pixels of a window live in lanes l, l^1, l^8, l^9).
//
// CTA pairs (template parameter CTA2): clusters of two CTAs issue every MMA together
// (tcgen05.mma.cta_group::2, M = 256), each CTA staging half of the weight rows.
#pragma once
#include <cuda_bf16.h>

#include "bias_mma.cuh"
#include "epilogue.cuh"
#include "gsd_ptx.cuh"

namespace gsd {

constexpr int kHaloRows = 18 * 10;                 // halo pixels per tile

struct HaloParams {
  CUtensorMap tm_src0;   // (C0, W, H, B) bf16, box (BKB/2, 10, 18, 1)
  CUtensorMap tm_src1;   // second source of the virtual concat
  CUtensorMap tm_w;      // (9*(C0+C1), Cout) bf16, box (BKB/2, BN)
  const float* scale;    // [Cout]
  const float* shift;    // [Cout]
  const float* bias;     // [Cout] or null: additive constant applied by one extra UMMA per tile (bias_mma.cuh) instead of
                         //   the epilogue; needs Cout == BN (a single N tile); scale / shift are then normally null
  __nv_bfloat16* out;    // (B, H, W, Cout) or null (head-only)
  __nv_bfloat16* out2;   // split output (dgrad of the decoder's concat conv, unet.py:48): channels >= split_c go to the dense
  int split_c;           //   tensor out2 (B, H, W, Cout - split_c), channels < split_c to out (B, H, W, split_c); null: one tensor
  __nv_bfloat16* pooled; // (B, H/2, W/2, Cout) or null
  // fused OutConv 1x1 + bias + depth de-normalisation (unet.py:54, normalization_utils.py:129); Cout == BN == 64
  const float* head_w;   // [ncls][64] or null
  const float* head_b;   // [ncls]
  float* head_y;         // (B, ncls, H, W) fp32 NCHW
  float* stats;          // global [2][Cout] (sum, sum of squares of the raw conv output) or null -- train-mode BatchNorm
  float head_scale, head_shift;
  int head_ncls;
  int cb0, cb1;          // channel blocks (BKB/2 channels each) of source 0 / 1
  int off_x, off_y;      // F.pad left/top of source 1
  int tiles_x, tiles_y, batch;
  int H, W, Cout;
  int n_tiles;           // Cout / BN
  int relu;
  int na, nb;            // ring depths (A halo buffers, B weight stages)
  FastDiv fd_ntiles, fd_tx, fd_ty;   // item -> (n tile, m group), m tile -> (tx, ty, b) without integer division
};

#ifndef GSD_ACC_BUFFERS_N64
#define GSD_ACC_BUFFERS_N64 4
#endif
constexpr int kAccBuffersN64 = GSD_ACC_BUFFERS_N64;

template <int BKB>
struct HaloGeom {
  static constexpr int BOX_BYTES = kHaloRows * BKB;                       // bytes one TMA halo box delivers
  static constexpr int BUF_BYTES = (BOX_BYTES + 1023) / 1024 * 1024;      // ring pitch
  static constexpr uint32_t LAYOUT = (BKB == 128) ? 2u : 6u;              // SWIZZLE_128B / SWIZZLE_32B
  // descriptor high words: SBO >> 4 @bits 32..45, version 1 @bit 46, swizzle @bits 61..63
  static constexpr uint32_t A_HI = ((10u * BKB) >> 4) | (1u << 14) | (LAYOUT << 29);   // image rows are 10 halo pixels apart
  static constexpr uint32_t B_HI = ((8u * BKB) >> 4) | (1u << 14) | (LAYOUT << 29);
  static constexpr int KSTEPS = BKB / 32;                                 // UMMA K = 16 bf16 = 32 bytes
  static constexpr int ROW16 = BKB / 16;                                  // one halo pixel in descriptor address units
};

// NEPI = epilogue warps (4: one per TMEM lane quadrant; 8: two sets of four that drain ALTERNATE work items, i.e.
// one set per TMEM accumulator buffer -- the per-tile bookkeeping is then paid once per tile, not once per unit)
// CTA2: thread-block cluster of two CTAs sharing every tcgen05.mma (cta_group::2, M = 256): each CTA stages and reads
// only HALF of the weight rows, so the smem operand traffic per UMMA drops from 32 + N/4 to 32 + N/8 wavefronts and the
// weight TMA traffic halves.  The pair walks pairs of M groups (same N tile); only the leader issues MMAs.
// SPLIT: two dense output tensors (HaloParams::out2 / split_c) -- a separate instantiation so that the pointer bookkeeping
// costs the single-output kernels nothing
template <int BN, int MT, bool WRES, int BKB, int NEPI, bool CTA2 = false, bool SPLIT = false>
__global__ void __launch_bounds__(64 + 32 * NEPI, (BKB == 32) ? 2 : 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int kHaloThreads = 64 + 32 * NEPI;
  using G = HaloGeom<BKB>;
  constexpr int BN_CTA = CTA2 ? BN / 2 : BN;          // weight rows this CTA stages
  constexpr int B_BYTES = BN_CTA * BKB;
  constexpr uint32_t NCTA = CTA2 ? 2 : 1;
  constexpr int KEL = BKB / 2;
  // accumulator ring in TMEM: 2 buffers, 4 for the one-K-block N = 64 tiles (36 UMMAs ~ 2 000 cycles per tile: with two buffers
  // the epilogue's round trip -- commit, tcgen05.ld, stores, arrive; across the cluster for a CTA pair -- is exposed)
  constexpr int NACC = (BN == 64) ? ((kAccBuffersN64 * MT * BN <= 512) ? kAccBuffersN64 : 512 / (MT * BN)) : 2;
  constexpr int TMEM_COLS = (NACC * MT * BN <= 128) ? 128 : (NACC * MT * BN <= 256) ? 256 : 512;
  static_assert(NACC * MT * BN <= 512 && (NACC & (NACC - 1)) == 0, "accumulators exceed TMEM");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int cbt = p.cb0 + p.cb1;
  const int nkb = 9 * cbt;
  const int na = p.na, nb = WRES ? 0 : p.nb;
  const uint32_t s_a = smem_base;
  const uint32_t s_b = s_a + na * G::BUF_BYTES;                        // B ring, or the resident weights
  const uint32_t s_bias = s_b + (WRES ? nkb : nb) * B_BYTES;          // 1 KB aligned; present only with p.bias
  const uint32_t s_aux = s_bias + (p.bias ? kBiasOnesBytes + 64 * 32 : 0);
  float* g_scale = reinterpret_cast<float*>(smem_gen + (s_aux - smem_base));   // [Cout] (<= 1024)
  float* g_shift = g_scale + p.Cout;
  float* g_head = g_shift + p.Cout;                                    // [4][64] + [4]
  const uint32_t s_epi = s_aux + 2 * p.Cout * 4 + (4 * 64 + 4) * 4 + 16;   // per-warp 2 KB epilogue transpose patches
  const uint32_t s_st = s_epi + NEPI * kEpiStageBytesPerWarp;              // per-CTA partial batch statistics [2][Cout]
  float* g_stats = reinterpret_cast<float*>(smem_gen + (s_st - smem_base));
  const uint32_t s_bar = s_st + 2 * p.Cout * 4;
  const uint32_t bar_fullA = s_bar;                    // [na]
  const uint32_t bar_emptyA = bar_fullA + 8 * na;      // [na]
  const uint32_t bar_fullB = bar_emptyA + 8 * na;      // [nb] (or [1] = resident weights landed)
  const uint32_t bar_emptyB = bar_fullB + 8 * (WRES ? 1 : nb);
  const uint32_t bar_acc_full = bar_emptyB + 8 * (WRES ? 1 : nb);   // [NACC]
  const uint32_t bar_acc_empty = bar_acc_full + 8 * NACC;            // [NACC]
  const uint32_t s_tmem_slot = bar_acc_empty + 8 * NACC;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_src0);
    tma_prefetch_desc(&p.tm_w);
    if (p.cb1) tma_prefetch_desc(&p.tm_src1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < na; ++i) { mbar_init(bar_fullA + 8 * i, 1); mbar_init(bar_emptyA + 8 * i, 1); }
    for (int i = 0; i < (WRES ? 1 : nb); ++i) { mbar_init(bar_fullB + 8 * i, 1); mbar_init(bar_emptyB + 8 * i, 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, 4 * NCTA); }
    fence_barrier_init();
  }
  if (warp == 2) { if (CTA2) tmem_alloc_2sm<TMEM_COLS>(s_tmem_slot); else tmem_alloc<TMEM_COLS>(s_tmem_slot); }
  for (int i = threadIdx.x; i < p.Cout; i += kHaloThreads) {
    if (p.scale) g_scale[i] = __ldg(p.scale + i);      // null: scale == 1 / shift == 0, the epilogue skips the loads
    if (p.shift) g_shift[i] = __ldg(p.shift + i);
    g_stats[i] = 0.f;
    g_stats[p.Cout + i] = 0.f;
  }
  if (p.bias) {
    bias_mma_fill(s_bias, p.bias + rank * BN_CTA, BN_CTA, threadIdx.x, kHaloThreads);
    fence_proxy_async_smem();
  }
  if (p.head_w) {
    for (int i = threadIdx.x; i < p.head_ncls * 64; i += kHaloThreads) g_head[i] = __ldg(p.head_w + i);
    if (threadIdx.x < p.head_ncls) g_head[256 + threadIdx.x] = __ldg(p.head_b + threadIdx.x);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_launch_dependents();      // the next layer may start its own setup while this grid drains

  const int m_tiles = p.tiles_x * p.tiles_y * p.batch;
  const int m_groups = (m_tiles + MT - 1) / MT;
  const int total_items = (CTA2 ? (m_groups + 1) / 2 : m_groups) * p.n_tiles;   // CTA2: items are PAIRS of M groups
  const int first_item = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      if (WRES) {
        if (leader) mbar_arrive_expect_tx(bar_fullB, (uint32_t)(nkb * B_BYTES) * NCTA);
        for (int kb = 0; kb < nkb; ++kb) {
          if (CTA2) tma_load_2d_2sm(s_b + kb * B_BYTES, &p.tm_w, bar_fullB, kb * KEL, (int)rank * BN_CTA);
          else tma_load_2d(s_b + kb * B_BYTES, &p.tm_w, bar_fullB, kb * KEL, 0);
        }
      }
      int ia = 0, ib = 0;
      uint32_t pa = 0, pb = 0;
      pdl_wait();               // activations of the previous layer are complete from here on (weights were not its output)
      for (int item = first_item; item < total_items; item += item_stride) {
        uint32_t mg, nt;
        fdivmod((uint32_t)item, p.fd_ntiles, mg, nt);
        if (CTA2) mg = 2 * mg + rank;
        for (int cb = 0; cb < cbt; ++cb) {
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            int mt = mg * MT + j;
            if (mt >= m_tiles) mt = m_tiles - 1;      // dummy duplicate keeps the pipeline protocol uniform
            uint32_t row, tx, b, ty;
            fdivmod((uint32_t)mt, p.fd_tx, row, tx);
            fdivmod(row, p.fd_ty, b, ty);
            const int xs = tx * 8 - 1, ys = ty * 16 - 1;
            mbar_wait(bar_emptyA + 8 * ia, pa ^ 1);
            if (leader) mbar_arrive_expect_tx(bar_fullA + 8 * ia, G::BOX_BYTES * NCTA);   // both CTAs' boxes land on the leader's barrier
            const CUtensorMap* tm = (cb < p.cb0) ? &p.tm_src0 : &p.tm_src1;
            const int cc = (cb < p.cb0) ? cb * KEL : (cb - p.cb0) * KEL;
            const int cx = (cb < p.cb0) ? xs : xs - p.off_x, cy = (cb < p.cb0) ? ys : ys - p.off_y;
            if (CTA2) tma_load_4d_2sm(s_a + ia * G::BUF_BYTES, tm, bar_fullA + 8 * ia, cc, cx, cy, b);
            else tma_load_4d(s_a + ia * G::BUF_BYTES, tm, bar_fullA + 8 * ia, cc, cx, cy, b);
            if (++ia == na) { ia = 0; pa ^= 1; }
          }
          if (!WRES) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(bar_emptyB + 8 * ib, pb ^ 1);
              if (leader) mbar_arrive_expect_tx(bar_fullB + 8 * ib, B_BYTES * NCTA);
              if (CTA2) tma_load_2d_2sm(s_b + ib * B_BYTES, &p.tm_w, bar_fullB + 8 * ib, (tap * cbt + cb) * KEL, nt * BN + (int)rank * BN_CTA);
              else tma_load_2d(s_b + ib * B_BYTES, &p.tm_w, bar_fullB + 8 * ib, (tap * cbt + cb) * KEL, nt * BN);
              if (++ib == nb) { ib = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
   if (leader) {
    // ===================================================== MMA issuer (whole warp converged; one elected lane issues)
    constexpr uint32_t idesc = CTA2 ? make_idesc_bf16_m256(BN) : make_idesc_bf16_m128(BN);
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    if (WRES) { mbar_wait(bar_fullB, 0); tc_fence_after(); }
    int it = 0;
    for (int item = first_item; item < total_items; item += item_stride, ++it) {
      const int buf = it & (NACC - 1);
      mbar_wait(bar_acc_empty + 8 * buf, ((it / NACC) & 1) ^ 1);
      tc_fence_after();
      const bool bias_mma = p.bias != nullptr;
      if (bias_mma) {          // D = 1 x bias^T first; every tap then accumulates
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < MT; ++j) bias_mma_issue<CTA2>(tmem_base + (buf * MT + j) * BN, s_bias, idesc);
        }
        __syncwarp();
      }
      for (int cb = 0; cb < cbt; ++cb) {
        uint32_t a_lo[MT];     // low word of the halo descriptor at tap (0,0), k = 0
        int a_slot[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          mbar_wait(bar_fullA + 8 * ia, pa);
          a_lo[j] = ((s_a + ia * G::BUF_BYTES) & 0x3FFFFu) >> 4;
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
                // tap view = same halo buffer, start shifted by (ty3*10 + tx3) halo pixels; the hardware applies
                // the swizzle XOR on absolute smem address bits, so no base-offset is needed (verified on B200)
                const uint32_t a0 = a_lo[j] + (ty3 * 10 + tx3) * G::ROW16;
#pragma unroll
                for (int k = 0; k < G::KSTEPS; ++k) {
                  if (CTA2)
                    umma_bf16_lohi_2sm(d_tmem, (a0 + 2 * k) | (1u << 16), G::A_HI, (b_lo + 2 * k) | (1u << 16), G::B_HI, idesc,
                                       (k != 0) ? 1u : (uint32_t)(((cb | tap) != 0) || bias_mma));
                  else
                    umma_bf16_lohi(d_tmem, (a0 + 2 * k) | (1u << 16), G::A_HI, (b_lo + 2 * k) | (1u << 16), G::B_HI, idesc,
                                   (k != 0) ? 1u : (uint32_t)(((cb | tap) != 0) || bias_mma));
                }
              }
              if (!WRES) { if (CTA2) umma_commit_2sm(bar_emptyB + 8 * ib); else umma_commit(bar_emptyB + 8 * ib); }
            }
            __syncwarp();
            if (!WRES) { if (++ib == nb) { ib = 0; pb ^= 1; } }
          }
        }
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < MT; ++j) { if (CTA2) umma_commit_2sm(bar_emptyA + 8 * a_slot[j]); else umma_commit(bar_emptyA + 8 * a_slot[j]); }
        }
        __syncwarp();
      }
      if (elect_one()) { if (CTA2) umma_commit_2sm(bar_acc_full + 8 * buf); else umma_commit(bar_acc_full + 8 * buf); }
      __syncwarp();
    }
   }
  } else {
    // ===================================================== epilogue (warps 2..): quadrant q = warp % 4, set = (warp-2)/4
    const int q = warp & 3;
    const int eset = (NEPI == 8) ? ((warp - 2) >> 2) : 0;
    const int ew = warp - 2;
    const int ly = 4 * q + (lane >> 3), lx = lane & 7;
    const int Hp = p.H >> 1, Wp = p.W >> 1;
    const bool hx = lane & 1, hy = (lane >> 3) & 1;   // which half of a pooling exchange this lane keeps
    int it = 0;
    pdl_wait();                 // no global write before the predecessor grid has finished (it may still read our output buffers' neighbours)
    for (int item = first_item; item < total_items; item += item_stride, ++it) {
      const int buf = it & (NACC - 1);
      if (NEPI == 8 && (buf & 1) != eset) continue;    // the other warp set owns this accumulator buffer
      uint32_t mg, nt;
      fdivmod((uint32_t)item, p.fd_ntiles, mg, nt);
      if (CTA2) mg = 2 * mg + rank;
      mbar_wait(bar_acc_full + 8 * buf, (it / NACC) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int mt = mg * MT + j;
        if (mt >= m_tiles) break;           // warp-uniform
        uint32_t trow, tx, b, ty;
        fdivmod((uint32_t)mt, p.fd_tx, trow, tx);
        fdivmod(trow, p.fd_ty, b, ty);
        const int y = ty * 16 + ly, x = tx * 8 + lx;
        EpiPixel px;
        const bool valid = (y < p.H) && (x < p.W);
        px.store_out = p.out != nullptr;
        px.valid = valid;
        px.s_stats = p.stats ? g_stats : nullptr;
        px.stats_ch0 = nt * BN;
        px.stats_stride = p.Cout;
        px.pvalid = ((y >> 1) < Hp) && ((x >> 1) < Wp);
        px.hx = hx; px.hy = hy; px.ypart = 8;
        const int img_row0 = (int)b * p.H;              // 32-bit pixel arithmetic, one 64-bit multiply per pointer
        px.prow = p.pooled ? p.pooled + (size_t)(((int)b * Hp + (y >> 1)) * Wp + (x >> 1)) * p.Cout + nt * BN + (hx ? 16 : 0) + (hy ? 8 : 0)
                           : nullptr;
        const int row_c = SPLIT ? p.split_c : p.Cout;        // channels per pixel of `out`
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + (lane >> 2);
          const int yy = ty * 16 + 4 * q + (r >> 3), xx = tx * 8 + (r & 7);
          px.rp[i] = (p.out && yy < p.H && xx < p.W) ? p.out + (size_t)((img_row0 + yy) * p.W + xx) * row_c + nt * BN : nullptr;
        }
        const uint32_t t_row = tmem_base + (buf * MT + j) * BN + ((uint32_t)(q * 32) << 16);
        float hacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32)
        {
          if constexpr (SPLIT) {   // this 32-column unit belongs to `out` (channels < split_c) or to `out2`
            const bool second = (int)(nt * BN) + c0 >= p.split_c;       // warp-uniform
            __nv_bfloat16* const tbase = second ? p.out2 - p.split_c : p.out;     // "- split_c": "+ c0" then lands on the right channel
            const int pitch = second ? p.Cout - p.split_c : p.split_c;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 8 * i + (lane >> 2);
              const int yy = ty * 16 + 4 * q + (r >> 3), xx = tx * 8 + (r & 7);
              px.rp[i] = (yy < p.H && xx < p.W) ? tbase + (size_t)((img_row0 + yy) * p.W + xx) * pitch + nt * BN : nullptr;
            }
          }
          epilogue_32cols(t_row, c0, p.scale ? g_scale + nt * BN : nullptr, p.shift ? g_shift + nt * BN : nullptr, p.relu, px, s_epi + ew * kEpiStageBytesPerWarp, lane, hacc,
                          p.head_w ? g_head : nullptr, p.head_ncls);
        }
        if (p.head_w && valid) {
          const size_t plane = (size_t)p.H * p.W;
          float* yo = p.head_y + (size_t)b * p.head_ncls * plane + (size_t)y * p.W + x;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < p.head_ncls) yo[k * plane] = (hacc[k] + g_head[256 + k]) * p.head_scale + p.head_shift;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CTA2) mbar_arrive_leader(bar_acc_empty + 8 * buf); else mbar_arrive(bar_acc_empty + 8 * buf); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();      // neither CTA exits (or frees TMEM) while the other may still signal it
  if (p.stats) {
    for (int i = threadIdx.x; i < 2 * p.Cout; i += kHaloThreads) atomicAdd(p.stats + i, g_stats[i]);
  }
  if (warp == 2) { if (CTA2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base); else tmem_dealloc<TMEM_COLS>(tmem_base); }
}

}  // namespace gsd
