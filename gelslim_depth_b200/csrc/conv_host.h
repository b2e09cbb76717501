// Host side of the tcgen05 conv: tile-shape / BN selection, tensor-map construction, launch.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <mutex>
#include <vector>

#include "conv_tc.cuh"
#include "host_util.h"

namespace gsd {

struct ConvDesc {
  const void* src0 = nullptr; int C0 = 0;                    // (B,H,W,C0) bf16
  int C0_real = 0;             // channels that carry data (first layer: 3 / 6 of the 16 stored); 0 = C0.  FLOP accounting only
  const void* src1 = nullptr; int C1 = 0, H1 = 0, W1 = 0;    // (B,H1,W1,C1) bf16, placed at (off_y, off_x)
  int off_y = 0, off_x = 0;
  int B = 0, H = 0, W = 0;
  const void* w = nullptr;     // bf16 [groups*Cout][ntaps*(C0+C1)]
  int Cout = 0;                // per output view
  int groups = 1;              // 1: plain conv; 4: transposed-conv 2x2/s2 scatter
  int ntaps = 1;
  int8_t dy[kMaxTaps] = {0}, dx[kMaxTaps] = {0};
  const float* scale = nullptr; const float* shift = nullptr;   // [groups*Cout]
  int relu = 0;
  void* out = nullptr;         // (B,H,W,Cout) or, groups==4, (B,2H,2W,Cout)
  void* pooled = nullptr;      // (B,H/2,W/2,Cout) or null
  int block_n = 0;             // 0 = choose
  // fused OutConv 1x1 + bias + depth de-normalisation (halo kernel only, Cout == 64)
  const float* head_w = nullptr; const float* head_b = nullptr; float* head_y = nullptr;
  float head_scale = 1.f, head_shift = 0.f; int head_ncls = 0;
  float* stats = nullptr;      // [2][groups*Cout] fp32, accumulated (caller zeroes): sum / sum of squares of the raw output
  const float* bias = nullptr; // halo kernel, Cout == 64: additive constant applied by the tensor core (bias_mma.cuh)
  void* out2 = nullptr; int split_c = 0;   // halo kernel: channels >= split_c go to the dense tensor out2 (see HaloParams)
  // transposed-conv INPUT gradient: src0 is a dense (B, Hf, Wf, Cs) tensor holding the gradient of the (2H x 2W)
  // up-sampled map at offset (s2d_off_y, s2d_off_x); C0 must be 2*Cs, ntaps 2 (gy), the output domain is H x W.
  int s2d = 0, s2d_Hf = 0, s2d_Wf = 0, s2d_off_y = 0, s2d_off_x = 0;
};

// launch `kernel` with `params`, optionally as a programmatic dependent launch (see gsd_ptx.cuh)
template <class Params>
inline int launch_maybe_pdl(void (*kernel)(Params), const Params& params, int grid, int block, int smem, cudaStream_t st, int pdl,
                            int cluster = 1) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {           // thread-block cluster (CTA pair of one TPC for cta_group::2 MMAs)
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  GSD_CUDA(cudaLaunchKernelEx(&cfg, kernel, params));
  return 0;
}

struct ConvLaunch {
  ConvParams p;
  int pdl = 0;                 // programmatic dependent launch (inference plan only)
  int bn = 0, bkb = 0, grid = 0;
  int cta2 = 0;                // CTA pair per work item (cta_group::2)
  double flops = 0;            // 2*M*N*K of the real (unpadded) problem
};

inline void pick_tile(int H, int W, bool pool, int* th, int* tw) {
  static const int cand[5][2] = {{8, 16}, {16, 8}, {4, 32}, {2, 64}, {1, 128}};
  if (const char* e = getenv("GSD_FORCE_TILE")) {   // tuning experiments only: "THxTW"
    int a = 0, b = 0;
    if (sscanf(e, "%dx%d", &a, &b) == 2 && a * b == 128 && !(pool && b != 8 && b != 16)) { *th = a; *tw = b; return; }
  }
  long best = -1;
  for (auto& c : cand) {
    if (pool && c[1] != 8 && c[1] != 16) continue;   // the shuffle max-pool needs both window rows in one warp
    long covered = (long)((H + c[0] - 1) / c[0]) * c[0] * ((W + c[1] - 1) / c[1]) * c[1];
    if (best < 0 || covered < best) { best = covered; *th = c[0]; *tw = c[1]; }
  }
}

inline int build_conv_launch(const ConvDesc& d, int num_sms, ConvLaunch* L) {
  memset(L, 0, sizeof *L);
  GSD_CHECK(d.ntaps >= 1 && d.ntaps <= kMaxTaps, "conv: ntaps %d out of range", d.ntaps);
  GSD_CHECK(d.groups == 1 || d.groups == 4, "conv: groups must be 1 or 4");
  GSD_CHECK(d.Cout % 64 == 0 && d.Cout > 0, "conv: Cout %d must be a multiple of 64", d.Cout);
  int bkb;
  if (d.C0 % 64 == 0 && d.C1 % 64 == 0) bkb = 128;
  else if (d.C0 == 16 && d.C1 == 0) bkb = 32;
  else return fail(-1, "conv: unsupported input channels C0=%d C1=%d (need multiples of 64, or 16 for the first layer)", d.C0, d.C1);
  const int kel = bkb / 2;
  GSD_CHECK(!(d.pooled && d.groups != 1), "conv: pooling with the transposed-conv scatter is not supported");
  GSD_CHECK(!d.pooled || (d.H >= 2 && d.W >= 2), "conv: pooled output needs H,W >= 2");

  ConvParams& p = L->p;
  int th = 8, tw = 16;
  pick_tile(d.H, d.W, d.pooled != nullptr, &th, &tw);
  p.th = th; p.tw = tw;
  p.tiles_x = (d.W + tw - 1) / tw;
  p.tiles_y = (d.H + th - 1) / th;
  p.batch = d.B;
  const int m_tiles = p.tiles_x * p.tiles_y * d.B;
  const int ntot = d.groups * d.Cout;

  int bn = d.block_n;
  if (bn == 0) {
    bn = 64;
    const int cands[3] = {256, 128, 64};
    for (int c : cands) {
      // a tile may span several (dy,dx) groups of the transposed-conv scatter (A is then loaded once for all of them)
      if (d.groups == 1 ? (d.Cout % c != 0) : (ntot % c != 0 || (c % d.Cout != 0 && d.Cout % c != 0))) continue;
      if (bkb == 32 && c != 64) continue;
      if ((long)m_tiles * (ntot / c) >= 2L * num_sms || c == 64) { bn = c; break; }
    }
  }
  GSD_CHECK(bn == 64 || bn == 128 || bn == 256, "conv: block_n %d invalid", bn);
  GSD_CHECK(ntot % bn == 0 && (d.Cout % bn == 0 || (d.groups == 4 && bn % d.Cout == 0)),
            "conv: block_n %d incompatible with Cout %d (groups %d)", bn, d.Cout, d.groups);
  GSD_CHECK(!(bkb == 32 && bn != 64), "conv: first-layer path supports block_n 64 only");
  L->bn = bn; L->bkb = bkb;
  {
    // CTA pairs for the tensor-bound 3x3 layers (most transposed convs are HBM / epilogue-bound: no operand to save)
    const int mode = getenv("GSD_CTA2") ? atoi(getenv("GSD_CTA2")) : 1;
    const long pair_items = (long)((m_tiles + 1) / 2) * (ntot / bn);
    // transposed convs: only the K = 1024 one (up.0.up) is tensor-bound enough to gain as pairs (same-box A/B at batch 64 with
    // GSD_CTA2_CONVT = smallest paired K: up.0.up 0.166 -> 0.145 ms, up.1.up 0.187 -> 0.184, up.2.up 0.239 -> 0.264, up.3.up 0.382 -> 0.455)
    const int convt_k = getenv("GSD_CTA2_CONVT") ? atoi(getenv("GSD_CTA2_CONVT")) : 1024;
    const bool convt = mode >= 1 && d.groups == 4 && convt_k > 0 && d.C0 >= convt_k && pair_items >= num_sms / 2;
    L->cta2 = (bkb == 128 && num_sms % 2 == 0 && bn >= 128 && (d.groups == 1 || convt) &&
               (mode == 2 || convt || (mode == 1 && d.ntaps == 9 && pair_items >= num_sms / 2))) ? 1 : 0;
  }
  p.n_tiles = ntot / bn;
  p.cout_per_group = d.Cout;
  p.kb0 = d.C0 / kel;
  p.kb1 = d.C1 / kel;
  p.ntaps = d.ntaps;
  for (int i = 0; i < d.ntaps; ++i) { p.tap_dy[i] = d.dy[i]; p.tap_dx[i] = d.dx[i]; }
  p.off_x = d.off_x; p.off_y = d.off_y;
  p.scale = d.scale; p.shift = d.shift;
  p.relu = d.relu;
  p.do_pool = d.pooled ? 1 : 0;

  const CUtensorMapSwizzle swz = bkb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  p.src5 = d.s2d;
  if (d.s2d) {
    GSD_CHECK(d.ntaps == 2 && d.C1 == 0 && d.C0 % 128 == 0, "conv: space-to-depth source needs ntaps 2 and C0 = 2*Cs");
    const uint64_t Cs = d.C0 / 2;
    char* base = static_cast<char*>(const_cast<void*>(d.src0)) + ((uint64_t)d.s2d_off_y * d.s2d_Wf + d.s2d_off_x) * Cs * 2;
    uint64_t dims[5] = {(uint64_t)d.C0, (uint64_t)d.W, 2, (uint64_t)d.H, (uint64_t)d.B};
    uint64_t str[4] = {(uint64_t)d.C0 * 2, (uint64_t)d.s2d_Wf * Cs * 2, 2ull * d.s2d_Wf * Cs * 2, (uint64_t)d.s2d_Hf * d.s2d_Wf * Cs * 2};
    uint32_t box[5] = {(uint32_t)kel, (uint32_t)tw, 1, (uint32_t)th, 1};
    GSD_TRY(encode_bf16_map(&p.tm_src0, base, 5, dims, str, box, swz, false));
  } else {
    uint64_t dims[4] = {(uint64_t)d.C0, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.B};
    uint64_t str[3] = {(uint64_t)d.C0 * 2, (uint64_t)d.W * d.C0 * 2, (uint64_t)d.H * d.W * d.C0 * 2};
    uint32_t box[4] = {(uint32_t)kel, (uint32_t)tw, (uint32_t)th, 1};
    GSD_TRY(encode_bf16_map(&p.tm_src0, const_cast<void*>(d.src0), 4, dims, str, box, swz, false));
  }
  if (d.C1) {
    uint64_t dims[4] = {(uint64_t)d.C1, (uint64_t)d.W1, (uint64_t)d.H1, (uint64_t)d.B};
    uint64_t str[3] = {(uint64_t)d.C1 * 2, (uint64_t)d.W1 * d.C1 * 2, (uint64_t)d.H1 * d.W1 * d.C1 * 2};
    uint32_t box[4] = {(uint32_t)kel, (uint32_t)tw, (uint32_t)th, 1};
    GSD_TRY(encode_bf16_map(&p.tm_src1, const_cast<void*>(d.src1), 4, dims, str, box, swz, false));
  } else {
    p.tm_src1 = p.tm_src0;
  }
  {
    const uint64_t ktot = (uint64_t)d.ntaps * (d.C0 + d.C1);
    uint64_t dims[2] = {ktot, (uint64_t)ntot};
    uint64_t str[1] = {ktot * 2};
    uint32_t box[2] = {(uint32_t)kel, (uint32_t)(L->cta2 ? bn / 2 : bn)};
    GSD_TRY(encode_bf16_map(&p.tm_w, const_cast<void*>(d.w), 2, dims, str, box, swz, true));
  }
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.pooled = static_cast<__nv_bfloat16*>(d.pooled);
  p.H = d.H; p.W = d.W; p.groups = d.groups; p.ntot = ntot;
  p.fd_ntiles = make_fastdiv(p.n_tiles); p.fd_tx = make_fastdiv(p.tiles_x); p.fd_ty = make_fastdiv(p.tiles_y);
  p.fd_cpg = make_fastdiv(p.cout_per_group);
  GSD_CHECK(p.cout_per_group < 4096 && (long)d.B * d.H * d.W * d.groups < (1L << 31), "conv: output too large for 32-bit pixel arithmetic");
  GSD_CHECK((long)m_tiles * p.n_tiles < (1L << 24) && p.tiles_x < 4096 && p.tiles_y < 4096 && p.n_tiles < 4096, "conv: too many tiles for the 24-bit tile index");
  p.stats = d.stats;
  GSD_CHECK(ntot <= 2048, "conv: more than 2048 output channels per launch are not supported");
  const long total = (long)m_tiles * p.n_tiles;
  L->grid = (int)(total < num_sms ? total : num_sms);
  if (L->cta2) {
    const long pairs_total = (long)((m_tiles + 1) / 2) * p.n_tiles;
    L->grid = (int)(2 * (pairs_total < num_sms / 2 ? pairs_total : num_sms / 2));
  }
  L->flops = 2.0 * d.B * d.H * d.W * (double)ntot * d.ntaps * ((d.C0_real ? d.C0_real : d.C0) + d.C1);   // the real (unpadded) problem
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember, per device of this
// process, the largest value already set (a process normally drives one GPU, but nothing here depends on it).
struct SmemAttrCache {
  int set[64] = {0};
  template <typename K>
  int ensure(K kernel, int smem) {
    int dev = 0;
    GSD_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (set[dev] < smem) {
      GSD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set[dev] = smem;
    }
    return 0;
  }
};

template <int BN, int BKB>
inline int launch_conv_cfg(const ConvLaunch& L, cudaStream_t st) {
  if constexpr (BKB == 128 && BN >= 128) {
    if (L.cta2) {
      using Cfg2 = ConvCfg<BN, BKB, true>;
      static SmemAttrCache attr_cache2;
      GSD_TRY(attr_cache2.ensure(conv_tc_kernel<BN, BKB, true>, Cfg2::SMEM_BYTES));
      return launch_maybe_pdl(conv_tc_kernel<BN, BKB, true>, L.p, L.grid, kConvThreads, Cfg2::SMEM_BYTES, st, L.pdl, 2);
    }
  }
  using Cfg = ConvCfg<BN, BKB>;
  static SmemAttrCache attr_cache;
  GSD_TRY(attr_cache.ensure(conv_tc_kernel<BN, BKB>, Cfg::SMEM_BYTES));
  return launch_maybe_pdl(conv_tc_kernel<BN, BKB>, L.p, L.grid, kConvThreads, Cfg::SMEM_BYTES, st, L.pdl);
}

inline int run_conv_launch(const ConvLaunch& L, cudaStream_t st) {
  if (L.bkb == 32) return launch_conv_cfg<64, 32>(L, st);
  switch (L.bn) {
    case 64: return launch_conv_cfg<64, 128>(L, st);
    case 128: return launch_conv_cfg<128, 128>(L, st);
    case 256: return launch_conv_cfg<256, 128>(L, st);
  }
  return fail(-1, "conv: no kernel for block_n %d", L.bn);
}

}  // namespace gsd

// ------------------------------------------------------------------------------------------------
// Halo kernel (conv_halo.cuh) host side
#include "conv_halo.cuh"

namespace gsd {

struct HaloLaunch {
  HaloParams p;
  int pdl = 0;
  int bn = 0, mt = 0, wres = 0, bkb = 0, nepi = 8, grid = 0, smem = 0;
  int cta2 = 0;               // CTA pair per work item (cta_group::2 MMAs, half of the weight rows per CTA)
  double flops = 0;
};

// fraction of MMA rows that are real pixels with 16x8 tiles
inline double halo_tile_efficiency(int H, int W) {
  return (double)H * W / ((double)((H + 15) / 16) * 16 * ((W + 7) / 8) * 8);
}

inline bool halo_supported(const ConvDesc& d);

// Halo-resident kernel or tap-streaming kernel?  The fixed 16x8 tiling of the halo kernel wastes MMA rows on small
// images (51 % useful at 20x26 or 10x13), so large batches use the tap-streaming kernel there (free choice of tile
// shape).  When the whole layer is a single wave of CTAs anyway (batch 1-2: latency, not throughput), each CTA's time is
// set by the L2->SM fabric, and the halo kernel moves 2.3x fewer bytes per UMMA (one halo box per 9 taps): measured
// 41 -> 20 us for the 1024-channel bottleneck conv at batch 1.
inline bool prefer_halo(const ConvDesc& d, int num_sms) {
  if (getenv("GSD_NO_HALO") || !halo_supported(d)) return false;
  const long mtl = (long)((d.W + 7) / 8) * ((d.H + 15) / 16) * d.B;
  const bool single_wave = d.Cout >= 128 && mtl * (d.Cout / 128) <= num_sms;
  return (double)d.H * d.W / ((double)((d.H + 15) / 16) * 16 * ((d.W + 7) / 8) * 8) >= 0.75 || single_wave;
}

inline bool halo_supported(const ConvDesc& d) {
  if (d.ntaps != 9 || d.groups != 1) return false;
  const bool first = (d.C0 == 16 && d.C1 == 0 && d.Cout == 64);
  if (!first && (d.C0 % 64 || d.C1 % 64)) return false;
  if (d.Cout % 64 || d.Cout > 1024) return false;
  if (d.Cout != 64 && d.Cout % 128) return false;
  for (int t = 0; t < 9; ++t)
    if (d.dy[t] != t / 3 - 1 || d.dx[t] != t % 3 - 1) return false;
  return true;
}

inline int build_halo_launch(const ConvDesc& d, int num_sms, HaloLaunch* L) {
  memset(L, 0, sizeof *L);
  GSD_CHECK(halo_supported(d), "halo conv: unsupported shape (C0=%d C1=%d Cout=%d taps=%d groups=%d)", d.C0, d.C1, d.Cout,
            d.ntaps, d.groups);
  HaloParams& p = L->p;
  const int bkb = (d.C0 == 16) ? 32 : 128;
  const int kel = bkb / 2;
  p.cb0 = d.C0 / kel; p.cb1 = d.C1 / kel;
  const int cbt = p.cb0 + p.cb1;
  p.off_x = d.off_x; p.off_y = d.off_y;
  p.tiles_x = (d.W + 7) / 8; p.tiles_y = (d.H + 15) / 16; p.batch = d.B;
  p.H = d.H; p.W = d.W; p.Cout = d.Cout;
  p.scale = d.scale; p.shift = d.shift; p.relu = d.relu;
  p.out = static_cast<__nv_bfloat16*>(d.out);
  p.pooled = static_cast<__nv_bfloat16*>(d.pooled);
  p.head_w = d.head_w; p.head_b = d.head_b; p.head_y = d.head_y;
  p.head_scale = d.head_scale; p.head_shift = d.head_shift; p.head_ncls = d.head_ncls;
  p.stats = d.stats;
  p.bias = d.bias;
  p.out2 = static_cast<__nv_bfloat16*>(d.out2); p.split_c = d.out2 ? d.split_c : 0;
  GSD_CHECK(!d.out2 || (d.out && !d.pooled && !d.head_w && d.split_c > 0 && d.split_c < d.Cout && d.split_c % 32 == 0 && (d.Cout - d.split_c) % 8 == 0),
            "halo conv: split output needs 0 < split_c < Cout, split_c a multiple of 32, no pooling / head");
  GSD_CHECK(!d.bias || d.Cout == 64, "halo conv: the tensor-core bias needs Cout == 64 (one N tile)");
  GSD_CHECK(!d.head_w || (d.Cout == 64 && d.head_ncls >= 1 && d.head_ncls <= 4 && d.head_y && d.head_b),
            "halo conv: fused 1x1 head needs Cout == 64 and 1..4 classes");
  GSD_CHECK(d.out || d.head_w, "halo conv: no output requested");
  const int budget = 227 * 1024 - 1024;
  const int buf_bytes = (kHaloRows * bkb + 1023) / 1024 * 1024;
  int bn, mt, wres, nepi = 8;
  // 64 output channels: the whole weight matrix stays resident in smem for the life of the persistent CTA
  // (4 epilogue warps instead of 8 when that leaves too little room for the halo ring)
  // CTA pairs (cta_group::2, conv_halo.cuh): GSD_CTA2 = 1 (default) where measured faster, 0 never, 2 everywhere (tests)
  const int cta2_mode = getenv("GSD_CTA2") ? atoi(getenv("GSD_CTA2")) : 1;
  const long mtl_all = (long)((d.W + 7) / 8) * ((d.H + 15) / 16) * d.B;
  // 64 output channels with K >= 1152 (up.3.conv.0): as single CTAs the resident weights (144 KB) leave room for 4
  // epilogue warps only; a CTA pair with streamed weights (32 rows per CTA and stage) and M = 2 x 256 measured 7 % faster
  // (round 2: resident in a pair -- 72 KB per CTA, eight epilogue warps, 4-deep accumulator ring -- measured 1.309 vs 1.222 ms: streamed stays)
  const bool pair64 = cta2_mode == 1 && bkb == 128 && d.Cout == 64 && cbt >= 2 && num_sms % 2 == 0 &&
                      ((mtl_all + 1) / 2 + 1) / 2 >= num_sms / 2;
  if (d.Cout == 64 && !pair64 && 9 * cbt * 64 * bkb <= 150 * 1024 && !(bkb == 128 && getenv("GSD_WRES0"))) { bn = 64; mt = 1; wres = 1; nepi = (9 * cbt * 64 * bkb > 80 * 1024) ? 4 : 8; }
  else if (d.Cout == 64) { bn = 64; mt = 2; wres = 0; }
  else {
    // streamed weights: M = 256 (two tiles share every weight stage) x N = 128 when there is enough work to fill the
    // SMs; small batches trade weight re-use for parallelism (more, smaller work items)
    const long mtl = (long)((d.W + 7) / 8) * ((d.H + 15) / 16) * d.B;
    bn = 128; mt = 2; wres = 0;
    if (((mtl + 1) / 2) * (d.Cout / 128) < num_sms) mt = 1;
    bool small = false;
    if (mt == 1) {
      // small batches: pick N by the makespan of the grid, waves x measured cycles per UMMA (~62 for N = 64, ~98 for
      // N = 128, ~170 for N = 256): a smaller N means more work items, worth it only while they fit the same waves
      small = true;
      const int ns[3] = {64, 128, 256};
      const long cyc[3] = {62, 98, 170};
      long best = -1;
      for (int i = 0; i < 3; ++i) {
        if (d.Cout % ns[i]) continue;
        const long items_i = mtl * (d.Cout / ns[i]);
        const long t = ((items_i + num_sms - 1) / num_sms) * cyc[i];
        if (best < 0 || t < best) { best = t; bn = ns[i]; }
      }
    }
    // N = 256 (one tile, M = 128): one UMMA reads A (32 smem wavefronts) for 256 output channels -- 96 wavefronts per
    // 128 tensor cycles instead of 64 per 64, which leaves the smem pipe room for the epilogue.  Measured on B200
    // (tools/exp_bn256.py, batch 64): +7..10 % for 256->512, 512->256, 1024->512, +2 % for 512->512, no gain for
    // 256->256 and 128->256, hence the rule below.
    const bool big = (d.C0 + d.C1 >= 512) || d.Cout >= 512;
    const bool want256 = d.block_n == 256 || (d.block_n == 0 && (big || getenv("GSD_BN256_ALL")) && !getenv("GSD_NO_BN256"));
    if (!small && want256 && d.Cout % 256 == 0 && mtl * (d.Cout / 256) >= num_sms) { bn = 256; mt = 1; }
    if (d.block_n == 256 && d.Cout % 256 == 0) { bn = 256; mt = 1; }
  }
  GSD_CHECK(bkb == 128 || wres, "halo conv: first-layer path needs resident weights");
  // CTA pairs (cta_group::2).  GSD_CTA2: 1 = the measured rule below (default), 0 = never, 2 = every 64-channel-block layer (tests).
  // Measured on B200, batch 64 (tools/exp_cta2.py): streamed-weight layers gain 8-17 % as pairs.  Streamed-weight layers pair at
  // every batch size: at batch 1-3 the grid is one partial wave either way and a pair reads each weight stage once for two
  // M tiles (same-box A/B, tools/ab_forward.py: batch 1 0.362 -> 0.344 ms, batch 2 0.538 -> 0.506 ms, batch >= 4 unchanged).
  // Resident-weight N = 64 layers (inc.3, up.3.conv.3, their dgrads): slower as pairs with a 2-deep accumulator ring (the
  // epilogue's round trip across the cluster is exposed behind 36-UMMA tiles: inc.3 0.82 -> 1.04 ms), faster with the 4-deep ring
  // of conv_halo.cuh (same-box A/B at batch 64: inc.3 0.828 -> 0.784 ms, up.3.conv.3 0.828 -> 0.757; batch 16-32: 0-8 %; batch <= 8
  // no difference), hence pairs from ~100 tiles per SM on.
  const bool wres_pair = wres && bkb == 128 && d.Cout == 64 && mtl_all >= 100L * num_sms;
  L->cta2 = (bkb == 128 && num_sms % 2 == 0 &&
             (cta2_mode == 2 || (cta2_mode == 1 && ((!wres && mtl_all >= 2) || pair64 || wres_pair)))) ? 1 : 0;
  const int aux = 4 * d.Cout * 4 + 2048 + nepi * kEpiStageBytesPerWarp + 2048 + (d.bias ? kBiasOnesBytes + 64 * 32 : 0);
  const int b_bytes = (L->cta2 ? bn / 2 : bn) * bkb;
  if (wres) {
    p.nb = 0;
    p.na = (budget - aux - 9 * cbt * b_bytes) / buf_bytes;
    if (p.na > 6) p.na = 6;
    GSD_CHECK(p.na >= 2, "halo conv: resident weights leave no room for the halo ring");
    L->smem = p.na * buf_bytes + 9 * cbt * b_bytes + aux + 1024;
  } else {
    p.na = 4;
    int nb_max = 9;
    if (const char* e = getenv("GSD_NA")) p.na = atoi(e);            // tuning experiments only
    if (const char* e = getenv("GSD_NB_MAX")) nb_max = atoi(e);
    p.nb = (budget - aux - p.na * buf_bytes) / b_bytes;
    if (p.nb > nb_max) p.nb = nb_max;
    GSD_CHECK(p.nb >= 3, "halo conv: no room for the weight ring");
    L->smem = p.na * buf_bytes + p.nb * b_bytes + aux + 1024;
  }
  p.n_tiles = d.Cout / bn;
  p.fd_ntiles = make_fastdiv(p.n_tiles); p.fd_tx = make_fastdiv(p.tiles_x); p.fd_ty = make_fastdiv(p.tiles_y);
  GSD_CHECK((long)p.tiles_x * p.tiles_y * d.B * p.n_tiles < (1L << 24), "halo conv: too many tiles for the 24-bit tile index");
  L->bn = bn; L->mt = mt; L->wres = wres; L->bkb = bkb; L->nepi = nepi;
  const CUtensorMapSwizzle swz = bkb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  auto src_map = [&](CUtensorMap* m, const void* base, int C, int H, int W) -> int {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)d.B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)kel, 10, 18, 1};
    return encode_bf16_map(m, const_cast<void*>(base), 4, dims, str, box, swz, false);
  };
  GSD_TRY(src_map(&p.tm_src0, d.src0, d.C0, d.H, d.W));
  if (d.C1) GSD_TRY(src_map(&p.tm_src1, d.src1, d.C1, d.H1, d.W1));
  else p.tm_src1 = p.tm_src0;
  {
    const uint64_t ktot = 9ull * (d.C0 + d.C1);
    uint64_t dims[2] = {ktot, (uint64_t)d.Cout};
    uint64_t str[1] = {ktot * 2};
    uint32_t box[2] = {(uint32_t)kel, (uint32_t)(L->cta2 ? bn / 2 : bn)};
    GSD_TRY(encode_bf16_map(&p.tm_w, const_cast<void*>(d.w), 2, dims, str, box, swz, true));
  }
  const long m_tiles = (long)p.tiles_x * p.tiles_y * d.B;
  const long items = ((m_tiles + mt - 1) / mt) * p.n_tiles;
  if (getenv("GSD_DEBUG_LAUNCH"))
    fprintf(stderr, "halo conv %dx%d B=%d C=%d+%d->%d: bn=%d mt=%d wres=%d nepi=%d cta2=%d na=%d nb=%d smem=%d\n", d.H, d.W, d.B, d.C0,
            d.C1, d.Cout, bn, mt, wres, nepi, L->cta2, p.na, p.nb, L->smem);
  if (L->cta2) {
    const long pair_items = (((m_tiles + mt - 1) / mt + 1) / 2) * p.n_tiles;
    const long pairs = pair_items < num_sms / 2 ? pair_items : num_sms / 2;
    L->grid = (int)(2 * pairs);
    L->flops = 2.0 * d.B * d.H * d.W * (double)d.Cout * 9 * ((d.C0_real ? d.C0_real : d.C0) + d.C1);
    return 0;
  }
  // the first-layer variant (16-channel rows) is epilogue-latency-bound: two co-resident CTAs per SM (78 KB smem, 96
  // registers, 128 TMEM columns each) double the warps the schedulers can pick from
  const long ctas = (long)num_sms * (bkb == 32 ? 2 : 1);
  L->grid = (int)(items < ctas ? items : ctas);
  L->flops = 2.0 * d.B * d.H * d.W * (double)d.Cout * 9 * ((d.C0_real ? d.C0_real : d.C0) + d.C1);
  return 0;
}

template <int BN, int MT, bool WRES, int BKB, int NEPI>
inline int launch_halo_cfg(const HaloLaunch& L, cudaStream_t st) {
  if constexpr (BKB == 128 && !WRES && BN >= 128) {        // split output (dgrad of the concat convs): N >= 128, streamed weights
    if (L.p.out2) {
      if (L.cta2) {
        static SmemAttrCache attr_cache_s2;
        GSD_TRY(attr_cache_s2.ensure(conv_halo_kernel<BN, MT, WRES, BKB, NEPI, true, true>, L.smem));
        return launch_maybe_pdl(conv_halo_kernel<BN, MT, WRES, BKB, NEPI, true, true>, L.p, L.grid, 64 + 32 * NEPI, L.smem, st, L.pdl, 2);
      }
      static SmemAttrCache attr_cache_s;
      GSD_TRY(attr_cache_s.ensure(conv_halo_kernel<BN, MT, WRES, BKB, NEPI, false, true>, L.smem));
      return launch_maybe_pdl(conv_halo_kernel<BN, MT, WRES, BKB, NEPI, false, true>, L.p, L.grid, 64 + 32 * NEPI, L.smem, st, L.pdl);
    }
  }
  GSD_CHECK(!L.p.out2, "halo conv: split output is built for N >= 128 with streamed weights only");
  if constexpr (BKB == 128) {
    if (L.cta2) {
      static SmemAttrCache attr_cache2;
      GSD_TRY(attr_cache2.ensure(conv_halo_kernel<BN, MT, WRES, BKB, NEPI, true>, L.smem));
      return launch_maybe_pdl(conv_halo_kernel<BN, MT, WRES, BKB, NEPI, true>, L.p, L.grid, 64 + 32 * NEPI, L.smem, st, L.pdl, 2);
    }
  }
  static SmemAttrCache attr_cache;
  GSD_TRY(attr_cache.ensure(conv_halo_kernel<BN, MT, WRES, BKB, NEPI>, L.smem));
  return launch_maybe_pdl(conv_halo_kernel<BN, MT, WRES, BKB, NEPI>, L.p, L.grid, 64 + 32 * NEPI, L.smem, st, L.pdl);
}

inline int run_halo_launch(const HaloLaunch& L, cudaStream_t st) {
  if (L.bkb == 32) return launch_halo_cfg<64, 1, true, 32, 8>(L, st);
  if (L.bn == 64 && L.mt == 1 && L.wres && L.nepi == 8) return launch_halo_cfg<64, 1, true, 128, 8>(L, st);
  if (L.bn == 64 && L.mt == 1 && L.wres && L.nepi == 4) return launch_halo_cfg<64, 1, true, 128, 4>(L, st);
  if (L.bn == 64 && L.mt == 2 && !L.wres) return launch_halo_cfg<64, 2, false, 128, 8>(L, st);
  if (L.bn == 64 && L.mt == 1 && !L.wres) return launch_halo_cfg<64, 1, false, 128, 8>(L, st);
  if (L.bn == 128 && L.mt == 1 && !L.wres) return launch_halo_cfg<128, 1, false, 128, 8>(L, st);
  if (L.bn == 128 && L.mt == 2 && !L.wres) return launch_halo_cfg<128, 2, false, 128, 8>(L, st);
  if (L.bn == 256 && L.mt == 1 && !L.wres) return launch_halo_cfg<256, 1, false, 128, 8>(L, st);
  return fail(-1, "halo conv: no kernel for bn=%d mt=%d wres=%d nepi=%d", L.bn, L.mt, L.wres, L.nepi);
}

}  // namespace gsd

// ------------------------------------------------------------------------------------------------
// First conv with the fused input prologue (conv_first.cuh) host side
#include "conv_first.cuh"

namespace gsd {

constexpr int kFirstNPRO = 7, kFirstNSET = 4;

struct FirstLaunch {
  FirstParams p;
  int grid = 0, smem = 0, cmax = 8;
  double flops = 0;
};

// w: bf16 [64][9][16] (the TMA-fed first layer's operand), out: (B,H,W,64) bf16.  `pre` is filled per call.
inline int build_first_launch(const void* w, const float* bias, void* out, int B, int H, int W, int cin_real, int num_sms, FirstLaunch* L) {
  memset(L, 0, sizeof *L);
  FirstParams& p = L->p;
  p.bias = bias; p.out = static_cast<__nv_bfloat16*>(out);
  p.tiles_x = (W + 7) / 8; p.tiles_y = (H + 15) / 16; p.batch = B; p.H = H; p.W = W;
  p.relu = 1;
  p.na = 8;
  if (const char* e = getenv("GSD_FIRST_NA")) p.na = atoi(e);        // tuning experiments only
  p.fd_tx = make_fastdiv(p.tiles_x); p.fd_ty = make_fastdiv(p.tiles_y);
  const long items = (long)p.tiles_x * p.tiles_y * B;
  GSD_CHECK(items < (1L << 24) && p.tiles_x < 4096 && p.tiles_y < 4096, "first conv: too many tiles for the 24-bit tile index");
  GSD_CHECK((long)B * H * W < (1L << 31), "first conv: output too large for 32-bit pixel arithmetic");
  {
    uint64_t dims[2] = {9ull * 16, 64};
    uint64_t str[1] = {9ull * 16 * 2};
    uint32_t box[2] = {16, 64};
    GSD_TRY(encode_bf16_map(&p.tm_w, const_cast<void*>(w), 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B, true));
  }
  L->grid = (int)(items < num_sms ? items : num_sms);
  L->smem = p.na * kFirstBuf + kFirstWBytes + kBiasOnesBytes + 64 * 32 + 4 * kFirstNSET * kEpiStageBytesPerWarp + 16 * p.na +
            16 * kFirstNSET + 64 + 1024;
  L->flops = 2.0 * B * H * W * 64.0 * 9 * cin_real;
  L->cmax = cin_real;
  GSD_CHECK(cin_real == 3 || cin_real == 6, "first conv: the fused prologue is built for 3 or 6 input channels");
  return 0;
}

template <int C, int KIND>
inline int launch_first_cfg(const FirstLaunch& L, cudaStream_t st) {
  static SmemAttrCache attr_cache;
  GSD_TRY(attr_cache.ensure(conv_first_kernel<kFirstNPRO, kFirstNSET, C, KIND>, L.smem));
  // never a programmatic dependent launch: it is the first kernel of a forward and must see completed uploads
  return launch_maybe_pdl(conv_first_kernel<kFirstNPRO, kFirstNSET, C, KIND>, L.p, L.grid, 32 * (1 + kFirstNPRO + 4 * kFirstNSET), L.smem, st, 0);
}

inline int run_first_launch(const FirstLaunch& L, cudaStream_t st) {
  const int kind = L.p.pre.input_u8;
  if (L.cmax == 3) return kind == 0 ? launch_first_cfg<3, 0>(L, st) : kind == 1 ? launch_first_cfg<3, 1>(L, st) : launch_first_cfg<3, 2>(L, st);
  return kind == 0 ? launch_first_cfg<6, 0>(L, st) : kind == 1 ? launch_first_cfg<6, 1>(L, st) : launch_first_cfg<6, 2>(L, st);
}

}  // namespace gsd

// ------------------------------------------------------------------------------------------------
// wgrad kernel (wgrad_tc.cuh) host side
#include "wgrad_tc.cuh"

namespace gsd {

struct WgradLaunch {
  WgradParams p;
  int grid = 0, smem = 0, xb = 128;
  double flops = 0;
};

// Split-K plan of the XB = 128 weight-gradient kernel.  A dW block (128 co x 64 ci x 9 taps) is produced by n0 CTAs
// that each accumulate filter rows {0,1} over 1/n0 of the pixel tiles (2 row-passes per tile) and n1 CTAs that
// accumulate filter row {2} over 1/n1 of them.  CTAs are launched longest-first and the hardware block scheduler hands
// the next one to whichever SM frees up, so the cost of a plan is the makespan of that greedy schedule; every CTA
// also pays a fixed set-up + drain cost (`ovh`, in row-passes).  Exhaustive search over (n0, n1), grid <= 4 waves.
inline double wgrad_makespan(long heavy, double c0, long light, double c1, int P) {
  const long r = heavy / P, q = heavy % P;
  const double t_early = r * c0, t_late = (r + 1) * c0;      // P-q SMs are free at t_early, q at t_late
  const double heavy_end = q ? t_late : t_early;
  if (light == 0) return heavy_end;
  auto fits = [&](double T) {
    long n = 0;
    if (T > t_early) n += (P - q) * (long)((T - t_early) / c1 + 1e-9);
    if (T > t_late) n += q * (long)((T - t_late) / c1 + 1e-9);
    return n >= light;
  };
  double lo = t_early, hi = t_late + (double)((light + P - 1) / P + 1) * c1;
  for (int it = 0; it < 60; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (fits(mid)) hi = mid; else lo = mid;
  }
  return hi > heavy_end ? hi : heavy_end;
}

inline void wgrad_choose_split(int blocks, long m_tiles, int num_sms, int* n0_out, int* n1_out) {
  if (const char* e = getenv("GSD_WG_SPLIT")) {            // experiments: "n0,n1"
    int a = 0, b = 0;
    if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0) { *n0_out = a; *n1_out = b; return; }
  }
  struct Key { int blocks; long m_tiles; int sms; int n0, n1; };
  static std::mutex mu;
  static std::vector<Key> cache;                           // the search costs ~10 ms: once per layer shape
  {
    std::lock_guard<std::mutex> g(mu);
    for (const Key& k : cache)
      if (k.blocks == blocks && k.m_tiles == m_tiles && k.sms == num_sms) { *n0_out = k.n0; *n1_out = k.n1; return; }
  }
  double ovh = 12.0, lf = 1.0;
  if (const char* e = getenv("GSD_WG_OVH")) ovh = atof(e);
  if (const char* e = getenv("GSD_WG_LF")) lf = atof(e);
  const long max_grid = 4L * num_sms;
  double best = 1e300;
  int bn0 = 1, bn1 = 1;
  for (int n0 = 1; n0 <= m_tiles && (long)blocks * (n0 + 1) <= (max_grid > blocks * 2L ? max_grid : blocks * 2L); ++n0) {
    const double c0 = 2.0 * (double)((m_tiles + n0 - 1) / n0) + ovh;
    for (int n1 = 1; n1 <= n0 && (long)blocks * (n0 + n1) <= (max_grid > blocks * 2L ? max_grid : blocks * 2L); ++n1) {
      const double c1 = lf * (double)((m_tiles + n1 - 1) / n1) + ovh;
      // longest-first: the kernel launches the two-row items first; if the one-row items are the longer ones the
      // estimate is still an upper bound of the in-order greedy schedule within one item
      const double t = c0 >= c1 ? wgrad_makespan((long)blocks * n0, c0, (long)blocks * n1, c1, num_sms)
                                : wgrad_makespan((long)blocks * n1, c1, (long)blocks * n0, c0, num_sms);
      if (t < best - 1e-9) { best = t; bn0 = n0; bn1 = n1; }
    }
  }
  *n0_out = bn0; *n1_out = bn1;
  std::lock_guard<std::mutex> g(mu);
  cache.push_back(Key{blocks, m_tiles, num_sms, bn0, bn1});
}

inline int build_wgrad_launch(const void* x0, int C0, const void* x1, int C1, int H1, int W1, int off_y, int off_x,
                              const void* dz, int Cout, int B, int H, int W, float* dw, int num_sms, WgradLaunch* L) {
  memset(L, 0, sizeof *L);
  const bool first = (C0 == 16 && C1 == 0);          // the 16-channel padded network input: N = 16 per tap, SWIZZLE_32B
  GSD_CHECK(first || (C0 % 64 == 0 && C1 % 64 == 0 && C0 > 0), "wgrad: input channels must be multiples of 64, or 16 (C0=%d C1=%d)", C0, C1);
  GSD_CHECK(Cout == 64 || Cout % 128 == 0, "wgrad: Cout must be 64 or a multiple of 128");
  WgradParams& p = L->p;
  const int xb = first ? 32 : 128, nt = xb / 2;
  L->xb = xb;
  p.cb0 = C0 / nt; p.cb1 = C1 / nt;
  p.off_x = off_x; p.off_y = off_y;
  const bool half_m = !first && Cout == 64;          // two filter rows per UMMA through a one-row-shifted dZ half
  // pixel tile 8 x 16 or 16 x 8: whichever wastes fewer UMMA rows on this image size (20x26: 51 % -> 68 % useful)
  const int hs = H + (half_m ? 1 : 0);
  const long pad8 = (long)((W + 7) / 8 * 8) * ((hs + 15) / 16 * 16), pad16 = (long)((W + 15) / 16 * 16) * ((hs + 7) / 8 * 8);
  p.tw16 = (!first && pad16 * 21 < pad8 * 20 && !getenv("GSD_WG_NO_TW16")) ? 1 : 0;
  const int TW = p.tw16 ? 16 : 8, TH = 128 / TW;
  p.tiles_x = (W + TW - 1) / TW; p.tiles_y = (hs + TH - 1) / TH; p.batch = B;
  p.Cout = Cout; p.co_blocks = (Cout + 127) / 128;
  p.dw = dw;
  p.stages = 4;
  const int blocks = p.co_blocks * (p.cb0 + p.cb1);
  const long m_tiles = (long)p.tiles_x * p.tiles_y * B;
  p.blocks = blocks;
  if (first || half_m) {
    int split = blocks >= num_sms ? 1 : num_sms / blocks;   // grid <= #SMs: no nearly empty second wave
    if (split > m_tiles) split = (int)m_tiles;
    p.split = split < 1 ? 1 : split;
    p.n0 = p.split; p.n1 = 0;                               // Cout == 64: one CTA accumulates all nine taps
    L->grid = blocks * p.split;
  } else {
    wgrad_choose_split(blocks, m_tiles, num_sms, &p.n0, &p.n1);
    L->grid = blocks * (p.n0 + p.n1);
  }
  if (const char* e = getenv("GSD_WG_FLAGS")) p.flags = atoi(e);
  L->smem = p.stages * (2 * kWgDzBytes + (180 * xb + 1023) / 1024 * 1024) + 1024 + 512;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)(TH + (half_m ? 1 : 0)), 1};
    GSD_TRY(encode_bf16_map(&p.tm_dz, const_cast<void*>(dz), 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, false));
  }
  auto src_map = [&](CUtensorMap* m, const void* base, int C, int h, int w) -> int {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
    uint32_t box[4] = {(uint32_t)nt, (uint32_t)(TW + 2), (uint32_t)(TH + 2), 1};
    return encode_bf16_map(m, const_cast<void*>(base), 4, dims, str, box,
                           first ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, false);
  };
  GSD_TRY(src_map(&p.tm_x0, x0, C0, H, W));
  if (C1) GSD_TRY(src_map(&p.tm_x1, x1, C1, H1, W1));
  else p.tm_x1 = p.tm_x0;
  L->flops = 2.0 * B * H * W * (double)Cout * 9 * (C0 + C1);
  return 0;
}

template <int XB>
inline int launch_wgrad_cfg(const WgradLaunch& L, cudaStream_t st) {
  static SmemAttrCache attr_cache;
  GSD_TRY(attr_cache.ensure(wgrad_tc_kernel<XB>, L.smem));
  wgrad_tc_kernel<XB><<<L.grid, kWgThreads, L.smem, st>>>(L.p);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

inline int run_wgrad_launch(const WgradLaunch& L, cudaStream_t st) {
  return L.xb == 32 ? launch_wgrad_cfg<32>(L, st) : launch_wgrad_cfg<128>(L, st);
}

}  // namespace gsd

namespace gsd {

struct WgradPwLaunch {
  WgradPwParams p;
  int grid = 0, smem = 0;
};

// in: (B,H,W,Cin) bf16; du: dense (B,Hf,Wf,Cout) bf16 holding the gradient of the up-sampled (2H x 2W) map at
// offset (off_y, off_x); dw: (Cin, Cout, 2, 2) fp32 accumulated.
inline int build_wgrad_pw_launch(const void* in, int Cin, const void* du, int Cout, int Hf, int Wf, int off_y, int off_x,
                                 int B, int H, int W, float* dw, int num_sms, WgradPwLaunch* L) {
  memset(L, 0, sizeof *L);
  GSD_CHECK(Cin % 128 == 0 && Cout % 64 == 0, "convT wgrad: need Cin %% 128 == 0 and Cout %% 64 == 0 (Cin=%d Cout=%d)", Cin, Cout);
  WgradPwParams& p = L->p;
  p.tiles_x = (W + 7) / 8; p.tiles_y = (H + 15) / 16; p.batch = B;
  p.Cin = Cin; p.Cout = Cout; p.dw = dw;
  const int blocks = (Cin / 128) * (Cout / 64);
  const long m_tiles = (long)p.tiles_x * p.tiles_y * B;
  int split = blocks >= num_sms ? 1 : num_sms / blocks;   // grid <= #SMs: a second, nearly empty wave would double the time
  if (split > m_tiles) split = (int)m_tiles;
  if (split < 1) split = 1;
  p.split = split;
  L->grid = blocks * split;
  L->smem = 2 * kWpStageBytes + 1024 + 512;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {64, 8, 16, 1};
    GSD_TRY(encode_bf16_map(&p.tm_in, const_cast<void*>(in), 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, false));
  }
  for (int g = 0; g < 4; ++g) {
    const int gy = g >> 1, gx = g & 1;
    char* base = static_cast<char*>(const_cast<void*>(du)) + ((uint64_t)(off_y + gy) * Wf + off_x + gx) * Cout * 2;
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {2ull * Cout * 2, 2ull * Wf * Cout * 2, (uint64_t)Hf * Wf * Cout * 2};
    uint32_t box[4] = {64, 8, 16, 1};
    GSD_TRY(encode_bf16_map(&p.tm_du[g], base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, false));
  }
  return 0;
}

inline int run_wgrad_pw_launch(const WgradPwLaunch& L, cudaStream_t st) {
  static SmemAttrCache attr_cache;
  GSD_TRY(attr_cache.ensure(wgrad_pw_kernel, L.smem));
  wgrad_pw_kernel<<<L.grid, kWgThreads, L.smem, st>>>(L.p);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gsd
