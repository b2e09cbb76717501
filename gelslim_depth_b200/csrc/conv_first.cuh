// First conv of the network (inc.double_conv.0, unet.py:11) with the INPUT PROLOGUE fused into its load path:
// get_difference_image (image_utils.py:6-10), the Left/Right finger split (general_dataset.py:71),
// normalize_tactile_image (normalization_utils.py:29-34) and the NCHW fp32 / uint8 -> NHWC bf16 re-layout happen in
// the producer warps, which write the (16+2) x (8+2) pixel halo of a tile straight into shared memory in the layout
// the tensor core expects.  The normalised image never exists in HBM (the unfused path writes a 16-channel padded
// bf16 tensor and reads it back 1.4x: 64 B per pixel that the algorithm does not need).
//
// Same implicit GEMM as conv_halo.cuh's first-layer variant (M = 128 pixels of a 16 x 8 tile, N = 64, one K = 16
// UMMA per filter tap, 32-byte pixel rows, SWIZZLE_32B, shifted-descriptor tap views), but the A operand is produced
// by generic-proxy stores instead of TMA:
//   * pixel hp of the halo lives at byte hp*32; SWIZZLE_32B XORs address bit 4 with bit 7, so the 16-byte chunk with
//     channels 0..7 sits at hp*32 + 16*((hp>>2)&1) and the all-zero chunk (channels 8..15) in the other half -- the
//     zero halves are written once per CTA, every tile only rewrites the data halves (one 16-byte store per pixel);
//   * writers fence with fence.proxy.async and arrive (32 lanes) on the slot's "full" mbarrier; tcgen05.commit
//     releases the slot.
// One CTA per SM: 1 MMA warp, NPRO producer warps (each owns every NPRO-th tile, so their global-load latencies
// overlap), 4*NSET epilogue warps -- NSET sets of four, one set per TMEM accumulator buffer, exactly the warp count the
// two co-resident CTAs of the TMA variant have (the layer is epilogue-issue / HBM-write bound).
#pragma once
#include <cuda_bf16.h>

#include "bias_mma.cuh"
#include "elementwise.cuh"
#include "epilogue.cuh"
#include "gsd_ptx.cuh"

namespace gsd {

struct FirstParams {
  PreParams pre;          // raw frames, base image, per-channel affine; pre.H/W == pre.Hr/Wr (no resampling on this path)
  CUtensorMap tm_w;       // (9*16, 64) bf16, box (16, 64), SWIZZLE_32B
  const float* bias;      // [64] folded BatchNorm shift (the scale lives in the weights): added by ONE extra UMMA per tile
                          //   (ones x bias, bias_mma.cuh), so the epilogue loads no per-channel constants at all
  __nv_bfloat16* out;     // (B, H, W, 64)
  int tiles_x, tiles_y, batch, H, W;
  int relu;
  int na;                 // halo ring depth (>= NPRO)
  FastDiv fd_tx, fd_ty;
};

constexpr int kFirstBuf = 6144;            // 180 halo pixels x 32 B = 5760, ring pitch rounded to 1 KB
constexpr int kFirstWBytes = 9 * 64 * 32;  // resident weights: 9 taps x (64 rows x 32 B)

// C: input channels (3 or 6: the reference's geometries; other counts take the two-pass path) -- the producer keeps a
// whole tile's loads in flight in registers (2 * 6 * C per lane).  KIND: 0 fp32 NCHW, 1 uint8 NCHW, 2 uint8 NHWC frames.
template <int NPRO, int NSET, int C, int KIND>
__global__ void __launch_bounds__(32 * (1 + NPRO + 4 * NSET), 1) conv_first_kernel(const __grid_constant__ FirstParams p) {
  constexpr int kThreads = 32 * (1 + NPRO + 4 * NSET);
  constexpr int NEPI = 4 * NSET;
  constexpr int TMEM_COLS = (NSET * 64 <= 128) ? 128 : (NSET * 64 <= 256) ? 256 : 512;
  static_assert(NSET * 64 <= 512, "accumulators exceed TMEM");
  static_assert((1 + NPRO) % 4 == 0, "epilogue warps must start at a multiple of 4 (TMEM lane quadrant = warp % 4)");
  // descriptor high words (conv_halo.cuh HaloGeom<32>): SBO = 10 halo pixels, version 1, SWIZZLE_32B
  constexpr uint32_t A_HI = ((10u * 32u) >> 4) | (1u << 14) | (6u << 29);
  constexpr uint32_t B_HI = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int na = p.na;
  const uint32_t s_a = smem_base;
  const uint32_t s_b = s_a + na * kFirstBuf;
  const uint32_t s_bias = s_b + kFirstWBytes;           // 1 KB-aligned: ones [128][32 B] + bias rows [64][32 B]
  const uint32_t s_epi = s_bias + kBiasOnesBytes + 64 * 32;
  const uint32_t s_bar = s_epi + NEPI * kEpiStageBytesPerWarp;
  const uint32_t bar_fullA = s_bar;                      // [na]  32 producer lanes arrive
  const uint32_t bar_emptyA = bar_fullA + 8 * na;        // [na]  tcgen05.commit
  const uint32_t bar_fullB = bar_emptyA + 8 * na;        // weights landed
  const uint32_t bar_acc_full = bar_fullB + 8;           // [NSET]
  const uint32_t bar_acc_empty = bar_acc_full + 8 * NSET;  // [NSET] 4 epilogue warps arrive
  const uint32_t s_tmem_slot = bar_acc_empty + 8 * NSET;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_w);
    for (int i = 0; i < na; ++i) { mbar_init(bar_fullA + 8 * i, 32); mbar_init(bar_emptyA + 8 * i, 1); }
    mbar_init(bar_fullB, 1);
    for (int i = 0; i < NSET; ++i) { mbar_init(bar_acc_full + 8 * i, 1); mbar_init(bar_acc_empty + 8 * i, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(s_tmem_slot);
  // zero the whole halo ring once: the channel 8..15 halves stay zero for the life of the CTA
  for (int i = threadIdx.x; i < na * kFirstBuf / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0u, 0u, 0u, 0u);
  bias_mma_fill(s_bias, p.bias, 64, threadIdx.x, kThreads);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_launch_dependents();

  const int total_items = p.tiles_x * p.tiles_y * p.batch;

  if (warp == 0) {
    // ===================================================== MMA issuer (+ the one-off weight load)
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_fullB, (uint32_t)kFirstWBytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_b + tap * 2048, &p.tm_w, bar_fullB, tap * 16, 0);
    }
    __syncwarp();
    constexpr uint32_t idesc = make_idesc_bf16_m128(64);
    mbar_wait(bar_fullB, 0);
    tc_fence_after();
    int it = 0, slot = 0;
    uint32_t pa = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int buf = it % NSET;
      mbar_wait(bar_acc_empty + 8 * buf, ((it / NSET) & 1) ^ 1);
      mbar_wait(bar_fullA + 8 * slot, pa);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = ((s_a + slot * kFirstBuf) & 0x3FFFFu) >> 4;
        const uint32_t d_tmem = tmem_base + buf * 64;
        bias_mma_issue<false>(d_tmem, s_bias, idesc);          // D = 1 x bias^T, the taps accumulate on top
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a0 = a_lo + ((tap / 3) * 10 + (tap % 3)) * 2;     // shifted tap view of the same halo
          const uint32_t b0 = ((s_b + tap * 2048) & 0x3FFFFu) >> 4;
          umma_bf16_lohi(d_tmem, a0 | (1u << 16), A_HI, b0 | (1u << 16), B_HI, idesc, 1u);
        }
        umma_commit(bar_emptyA + 8 * slot);
        umma_commit(bar_acc_full + 8 * buf);
      }
      __syncwarp();
      if (++slot == na) { slot = 0; pa ^= 1; }
    }
  } else if (warp <= NPRO) {
    // ===================================================== producers: raw frames -> normalised bf16 halo in smem
    // Straight-line code: every load goes to a CLAMPED (always valid) address and out-of-image pixels are zeroed by a
    // select afterwards, so the 2 * 6 * C loads of a tile are all in flight together (one DRAM round trip per tile).
    const PreParams& q = p.pre;
    const int r3 = lane / 10, cx = lane - r3 * 10;
    const bool active = lane < 30;
    constexpr int ct1 = C;
    const int ct = q.split_fingers ? 2 * ct1 : ct1;
    const int frames = q.split_fingers ? (q.B >> 1) : q.B;
    const int plane = p.H * p.W;                       // < 2^31 / 64 (checked on the host)
    const bool has_base = q.use_diff != 0;
    float sc[C], sh[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { sc[c] = q.in_scale[c]; sh[c] = q.in_shift[c]; }
    int it = warp - 1, slot = (warp - 1) % na;
    uint32_t phase = 0;                                // parity of this warp's next wait on bar_emptyA[slot], flips per ring lap
    for (int item = blockIdx.x + (warp - 1) * gridDim.x; item < total_items; item += NPRO * gridDim.x, it += NPRO) {
      uint32_t row, tx, b, ty;
      fdivmod((uint32_t)item, p.fd_tx, row, tx);
      fdivmod(row, p.fd_ty, b, ty);
      const int xs = (int)tx * 8 - 1, ys = (int)ty * 16 - 1;
      int f = 0, n = (int)b;
      if (q.split_fingers) { f = (int)b / frames; n = (int)b - f * frames; }
      // channel 0 of this network sample inside the raw tensor / the base image
      const unsigned char* raw8 = static_cast<const unsigned char*>(q.x) +
                                  (KIND == 2 ? (long)n * plane * ct + f * C : ((long)n * ct + f * C) * plane);
      const float* rawf = static_cast<const float*>(q.x) + ((long)n * ct + f * C) * plane;
      const float* basep = q.base + ((long)(q.base_batch == 1 ? 0 : n) * ct + f * C) * plane;
      const int x = xs + cx;
      const bool xin = active && x >= 0 && x < p.W;
      const int xc = min(max(x, 0), p.W - 1);
      // C = 6: two half-tile batches (3 row groups = 36 loads in flight per lane), so that the producer fits the 80
      // registers a 768-thread CTA leaves per thread; C = 3: the whole tile at once
      constexpr int GPB = (C <= 3) ? 6 : 3;
      const uint32_t dst = s_a + slot * kFirstBuf;
#pragma unroll 1
      for (int g0 = 0; g0 < 6; g0 += GPB) {
        uint32_t tv[GPB][C];           // raw bits: fp32 value or the zero-extended camera byte
        float bv[GPB][C];
        bool inb[GPB];
#pragma unroll
        for (int g = 0; g < GPB; ++g) {
          const int y = ys + (g0 + g) * 3 + r3;
          inb[g] = xin && y >= 0 && y < p.H;
          const int pix = min(max(y, 0), p.H - 1) * p.W + xc;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if (KIND == 0) tv[g][c] = __float_as_uint(__ldg(rawf + c * plane + pix));
            else if (KIND == 1) tv[g][c] = __ldg(raw8 + c * plane + pix);
            else tv[g][c] = __ldg(raw8 + pix * ct + c);
            bv[g][c] = has_base ? __ldg(basep + c * plane + pix) : 0.f;
          }
        }
        // ring slot of local item `it`: it % na, tracked incrementally (this warp advances by NPRO slots per tile)
        if (g0 == 0) mbar_wait(bar_emptyA + 8 * slot, phase ^ 1);
#pragma unroll
        for (int g = 0; g < GPB; ++g) {
          float v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            v[c] = 0.f;
            if (c < C) {
              float t = (KIND == 0) ? __uint_as_float(tv[g][c]) : (float)tv[g][c];
              if (has_base) t = (t - bv[g][c] + 255.0f) * 0.5f;
              v[c] = inb[g] ? fmaf(sc[c], t, sh[c]) : 0.f;      // outside the image: the conv's zero padding
            }
          }
          if (active) {
            const int hp = ((g0 + g) * 3 + r3) * 10 + cx;
            const uint32_t addr = dst + hp * 32 + (((hp >> 2) & 1) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(v[0], v[1])),
                         "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
          }
        }
      }
      fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(bar_fullA + 8 * slot);
      slot += NPRO;
      if (slot >= na) { slot -= na; phase ^= 1; }
    }
  } else {
    // ===================================================== epilogue: set = accumulator buffer, q = TMEM lane quadrant
    const int ew = warp - 1 - NPRO;
    const int eset = ew >> 2;
    const int q = warp & 3;
    const int ly = 4 * q + (lane >> 3), lx = lane & 7;
    int it = 0;
    pdl_wait();
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      if (it % NSET != eset) continue;
      uint32_t trow, tx, b, ty;
      fdivmod((uint32_t)item, p.fd_tx, trow, tx);
      fdivmod(trow, p.fd_ty, b, ty);
      mbar_wait(bar_acc_full + 8 * eset, (it / NSET) & 1);
      tc_fence_after();
      const int y = ty * 16 + ly, x = tx * 8 + lx;
      EpiPixel px;
      px.store_out = true;
      px.valid = (y < p.H) && (x < p.W);
      px.s_stats = nullptr; px.stats_ch0 = 0; px.stats_stride = 0;
      px.pvalid = false; px.hx = false; px.hy = false; px.ypart = 8; px.prow = nullptr;
      const int img_row0 = (int)b * p.H;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2);
        const int yy = ty * 16 + 4 * q + (r >> 3), xx = tx * 8 + (r & 7);
        px.rp[i] = (yy < p.H && xx < p.W) ? p.out + (size_t)((img_row0 + yy) * p.W + xx) * 64 : nullptr;
      }
      const uint32_t t_row = tmem_base + eset * 64 + ((uint32_t)(q * 32) << 16);
      float hacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32)
        epilogue_32cols(t_row, c0, nullptr, nullptr, p.relu, px, s_epi + ew * kEpiStageBytesPerWarp, lane, hacc, nullptr, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * eset);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace gsd
