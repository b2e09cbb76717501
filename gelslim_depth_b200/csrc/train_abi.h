// C ABI of the training-step operators (included by plan.cu).  Each op is one kernel launch on `stream`.
#pragma once
#include "train_ops.cuh"

using namespace gsd;

static int check_dev(int device, const char* who) {
  int major = 0;
  GSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  GSD_CHECK(major == 10, "%s: device %d is not sm_100 (no fallback)", who, device);
  return 0;
}
static int num_sms_of(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms;
}

// Generic conv launch used by the training path: picks the halo-resident or the tap-streaming kernel exactly like
// the inference plan.  ntaps 9 = 3x3/pad 1, ntaps 1 = pointwise (groups 4: transposed-conv scatter).
extern "C" int gsd_op_conv_auto_bf16(const void* src0, int C0, const void* src1, int C1, int H1, int W1, int off_y,
                                     int off_x, int B, int H, int W, const void* w, int Cout, int ntaps, int groups,
                                     const float* scale, const float* shift, int relu, void* out, void* pooled,
                                     float* stats, int device, void* stream) {
  GSD_CHECK(src0 && w && out, "gsd_op_conv_auto_bf16: null argument");     // scale / shift may be NULL (= 1 / 0)
  GSD_DEVICE(device);
  GSD_TRY(check_dev(device, "gsd_op_conv_auto_bf16"));
  ConvDesc d;
  d.src0 = src0; d.C0 = C0; d.src1 = src1; d.C1 = src1 ? C1 : 0; d.H1 = H1; d.W1 = W1; d.off_y = off_y; d.off_x = off_x;
  d.B = B; d.H = H; d.W = W; d.w = w; d.Cout = Cout; d.groups = groups;
  if (ntaps == 9) taps3x3(&d);
  else { GSD_CHECK(ntaps == 1, "gsd_op_conv_auto_bf16: ntaps must be 9 or 1"); d.ntaps = 1; }
  d.scale = scale; d.shift = shift; d.relu = relu; d.out = out; d.pooled = pooled; d.stats = stats;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (prefer_halo(d, num_sms_of(device))) {
    HaloLaunch L;
    GSD_TRY(build_halo_launch(d, num_sms_of(device), &L));
    return run_halo_launch(L, st);
  }
  ConvLaunch L;
  GSD_TRY(build_conv_launch(d, num_sms_of(device), &L));
  return run_conv_launch(L, st);
}

// Launch planning without a GPU (tests/test_host_rules_cpu.py): which kernel and configuration a 3x3 conv layer gets.
// out[0] = 1 halo-resident / 0 tap-streaming kernel, [1] N per UMMA, [2] M tiles per item, [3] resident weights,
// [4] epilogue warps, [5] CTA pair, [6] halo ring, [7] weight ring, [8] dynamic smem bytes, [9] grid, [10] tile h, [11] tile w
extern "C" int gsd_debug_plan_conv3x3(int B, int H, int W, int C0, int C1, int Cout, int num_sms, int* out) {
  GSD_CHECK(out && B >= 1 && H >= 1 && W >= 1 && num_sms >= 1, "gsd_debug_plan_conv3x3: bad argument");
  ConvDesc d;
  void* dummy = reinterpret_cast<void*>(static_cast<uintptr_t>(256));
  d.src0 = dummy; d.C0 = C0; d.src1 = C1 ? dummy : nullptr; d.C1 = C1; d.H1 = H; d.W1 = W;
  d.B = B; d.H = H; d.W = W; d.w = dummy; d.Cout = Cout; d.groups = 1;
  taps3x3(&d);
  d.out = dummy;
  for (int i = 0; i < 12; ++i) out[i] = 0;
  plan_only_mode() = true;
  int rc = 0;
  if (prefer_halo(d, num_sms)) {
    HaloLaunch L;
    rc = build_halo_launch(d, num_sms, &L);
    if (rc == 0) {
      const int v[12] = {1, L.bn, L.mt, L.wres, L.nepi, L.cta2, L.p.na, L.p.nb, L.smem, L.grid, 16, 8};
      for (int i = 0; i < 12; ++i) out[i] = v[i];
    }
  } else {
    ConvLaunch L;
    rc = build_conv_launch(d, num_sms, &L);
    if (rc == 0) {
      const int v[12] = {0, L.bn, 1, 0, kEpiWarps, L.cta2, 0, 0, 0, L.grid, L.p.th, L.p.tw};
      for (int i = 0; i < 12; ++i) out[i] = v[i];
    }
  }
  plan_only_mode() = false;
  return rc;
}

// Transposed-conv input gradient: d_in[b,y,x,ci] = sum_{gy,gx,co} dU[b,2y+gy,2x+gx,co] Wt[ci,co,gy,gx]
// du: dense (B,Hf,Wf,Cs) bf16 with the up-sampled map's gradient at offset (off_y, off_x); w: bf16 [Cin][(gy,gx,co)].
extern "C" int gsd_op_convt_dgrad_bf16(const void* du, int Cs, int Hf, int Wf, int off_y, int off_x, const void* w, int Cin,
                                       int B, int H, int W, const float* scale, const float* shift, void* out, int device,
                                       void* stream) {
  GSD_CHECK(du && w && out, "gsd_op_convt_dgrad_bf16: null argument");        // scale / shift may be NULL (= 1 / 0)
  GSD_DEVICE(device);
  GSD_TRY(check_dev(device, "gsd_op_convt_dgrad_bf16"));
  ConvDesc d;
  d.src0 = du; d.C0 = 2 * Cs; d.B = B; d.H = H; d.W = W; d.w = w; d.Cout = Cin; d.groups = 1; d.ntaps = 2;
  d.s2d = 1; d.s2d_Hf = Hf; d.s2d_Wf = Wf; d.s2d_off_y = off_y; d.s2d_off_x = off_x;
  d.scale = scale; d.shift = shift; d.relu = 0; d.out = out;
  ConvLaunch L;
  GSD_TRY(build_conv_launch(d, num_sms_of(device), &L));
  return run_conv_launch(L, static_cast<cudaStream_t>(stream));
}

extern "C" int gsd_op_convt_wgrad_bf16(const void* in, int Cin, const void* du, int Cout, int Hf, int Wf, int off_y, int off_x,
                                       int B, int H, int W, float* dw, int device, void* stream) {
  GSD_CHECK(in && du && dw, "gsd_op_convt_wgrad_bf16: null argument");
  GSD_DEVICE(device);
  GSD_TRY(check_dev(device, "gsd_op_convt_wgrad_bf16"));
  WgradPwLaunch L;
  GSD_TRY(build_wgrad_pw_launch(in, Cin, du, Cout, Hf, Wf, off_y, off_x, B, H, W, dw, num_sms_of(device), &L));
  return run_wgrad_pw_launch(L, static_cast<cudaStream_t>(stream));
}

// (B,C,Hr,Wr) fp32 NCHW -> NHWC bf16 with channels padded to 16 (+ optional difference image / resample / affine)
extern "C" int gsd_op_prologue_bf16(const float* x, const float* base, int base_batch, int use_diff, int B, int Cc, int Hr,
                                    int Wr, int H, int W, const float* scale8_host, const float* shift8_host, void* out16,
                                    void* stream) {
  GSD_CHECK(x && out16 && scale8_host && shift8_host && Cc <= 8, "gsd_op_prologue_bf16: bad argument");
  GSD_DEVICE_OF(x);
  PreParams p;
  p.x = x; p.base = use_diff ? base : nullptr; p.base_batch = base_batch; p.use_diff = use_diff;
  p.B = B; p.C = Cc; p.Hr = Hr; p.Wr = Wr; p.H = H; p.W = W; p.split_fingers = 0; p.input_u8 = 0;
  for (int c = 0; c < 8; ++c) { p.in_scale[c] = scale8_host[c]; p.in_shift[c] = shift8_host[c]; }
  const int pg = ew_grid((long)B * H * ((W + 255) / 256) * 256, 256, 148 * 32);
  if (Hr == H && Wr == W) prologue_kernel<true><<<pg, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<__nv_bfloat16*>(out16));
  else prologue_kernel<false><<<pg, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, static_cast<__nv_bfloat16*>(out16));
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_negate_f32(const float* in, int n, float* out, void* stream) {
  GSD_CHECK(in && out && n > 0, "gsd_op_negate_f32: bad argument");
  GSD_DEVICE_OF(in);
  negate_f32_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(in, n, out);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_bn_finalize(const float* stats, double count, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, float momentum, float eps, int C, const float* neg_center, float* scale,
                                  float* shift, float* mean, float* rstd, long long* num_batches_tracked, void* stream) {
  GSD_CHECK(stats && gamma && beta && scale && shift && mean && rstd && C > 0, "gsd_op_bn_finalize: bad argument");
  GSD_DEVICE_OF(stats);
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, (float)count, gamma, beta, running_mean,
                                                                                    running_var, momentum, eps, C, neg_center, scale, shift,
                                                                                    mean, rstd, num_batches_tracked);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_bn_relu_apply(const void* z, const float* scale, const float* shift, int B, int H, int W, int C, void* a,
                                    void* pooled, void* stream) {
  GSD_CHECK(z && scale && shift && a && C % 8 == 0, "gsd_op_bn_relu_apply: bad argument");
  GSD_DEVICE_OF(z);
  const int C8 = C / 8;
  GSD_CHECK((C8 & (C8 - 1)) == 0 && C8 <= 256, "gsd_op_bn_relu_apply: C/8 must be a power of two <= 256");
  int c8_shift = 0;
  while ((1 << c8_shift) < C8) ++c8_shift;
  const long rows = (long)B * ((H + 1) / 2);
  const int grid = (int)(rows < 148 * 24 ? rows : 148 * 24);     // ~one window row per block iteration: measured best of 3..32
  bn_relu_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(z), scale, shift, B, H, W, C, c8_shift, static_cast<__nv_bfloat16*>(a),
      static_cast<__nv_bfloat16*>(pooled));
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// loss (1 float, accumulated: caller zeroes) and dy = 2 (y - t) / n   (train_unet.py:51-52,370)
extern "C" int gsd_op_mse(const float* y, const float* t, long long n, float* loss, float* dy, void* stream) {
  GSD_CHECK(y && t && loss && dy && n > 0, "gsd_op_mse: bad argument");
  GSD_DEVICE_OF(y);
  mse_kernel<<<ew_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, t, (long)n, loss, dy);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_head_fwd(const void* a, const float* w, const float* bias, int ncls, int B, int H, int W, float* y, void* stream) {
  GSD_CHECK(a && w && bias && y && ncls >= 1 && ncls <= 4, "gsd_op_head_fwd: bad argument");
  GSD_DEVICE_OF(a);
  const long npix = (long)H * W;
  head_kernel<64><<<ew_grid(npix * B), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(a), w, bias, ncls, 1.f,
                                                                                   0.f, npix, B, y);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// training forward of the last unit: y = OutConv(relu(BatchNorm(z))) without materialising the post-ReLU tensor
extern "C" int gsd_op_bn_relu_head_fwd(const void* z, const float* scale, const float* shift, const float* w, const float* bias,
                                       int ncls, int B, int H, int W, float* y, void* stream) {
  GSD_CHECK(z && scale && shift && w && bias && y && ncls >= 1 && ncls <= 4, "gsd_op_bn_relu_head_fwd: bad argument");
  GSD_DEVICE_OF(z);
  const long npix = (long)H * W;
  head_kernel<64><<<ew_grid(npix * B), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(z), w, bias, ncls, 1.f,
                                                                                   0.f, npix, B, y, scale, shift);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// backward of OutConv + ReLU + BatchNorm of the last unit in two passes over z (see head_bn_bwd_kernel)
template <int NCLS>
static int launch_head_bn_bwd(const __nv_bfloat16* z, const float* dy, const float* w, const float* scale, const float* shift,
                              const float* mean, const float* rstd, const float* gamma, float count, unsigned npix, unsigned total,
                              float* sums, float* dw, float* db, __nv_bfloat16* dz, cudaStream_t st, float* sums2 = nullptr) {
  long blocks = ((long)total * 8 + 255) / 256;
  const int grid = (int)(blocks < 148 * 2 ? blocks : 148 * 2);
  head_bn_bwd_kernel<NCLS, false><<<grid, 256, 0, st>>>(z, dy, w, scale, shift, mean, rstd, gamma, count, npix, total, sums, dw, db, dz, sums2);
  head_bn_bwd_kernel<NCLS, true><<<grid, 256, 0, st>>>(z, dy, w, scale, shift, mean, rstd, gamma, count, npix, total, sums, dw, db, dz, sums2);
  GSD_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int gsd_op_head_bn_bwd(const void* z, const float* dy, const float* w, const float* scale, const float* shift,
                                  const float* mean, const float* rstd, const float* gamma, double count, int ncls, int B, int H,
                                  int W, float* sums, float* dw, float* db, void* dz, void* stream) {
  GSD_CHECK(z && dy && w && scale && shift && mean && rstd && gamma && sums && dw && db && dz && ncls >= 1 && ncls <= 4,
            "gsd_op_head_bn_bwd: bad argument");
  GSD_DEVICE_OF(z);
  const long npix = (long)H * W;
  GSD_CHECK(npix * B < (1L << 31), "gsd_op_head_bn_bwd: more than 2^31 pixels");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* zb = static_cast<const __nv_bfloat16*>(z);
  __nv_bfloat16* dzb = static_cast<__nv_bfloat16*>(dz);
  const unsigned np = (unsigned)npix, tot = (unsigned)(npix * B);
  switch (ncls) {
    case 1: return launch_head_bn_bwd<1>(zb, dy, w, scale, shift, mean, rstd, gamma, (float)count, np, tot, sums, dw, db, dzb, st);
    case 2: return launch_head_bn_bwd<2>(zb, dy, w, scale, shift, mean, rstd, gamma, (float)count, np, tot, sums, dw, db, dzb, st);
    case 3: return launch_head_bn_bwd<3>(zb, dy, w, scale, shift, mean, rstd, gamma, (float)count, np, tot, sums, dw, db, dzb, st);
    default: return launch_head_bn_bwd<4>(zb, dy, w, scale, shift, mean, rstd, gamma, (float)count, np, tot, sums, dw, db, dzb, st);
  }
}

extern "C" int gsd_op_head_bwd(const void* a, const float* dy, const float* w, int ncls, int B, int H, int W, void* da, float* dw,
                               float* db, void* stream) {
  GSD_CHECK(a && dy && w && da && dw && db && ncls >= 1 && ncls <= 4, "gsd_op_head_bwd: bad argument");
  GSD_DEVICE_OF(a);
  const long npix = (long)H * W;
  GSD_CHECK(npix * B < (1L << 31), "gsd_op_head_bwd: more than 2^31 pixels");
  long blocks = (npix * B * 8 + 255) / 256;
  const int grid = (int)(blocks < 148 * 4 ? blocks : 148 * 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* ab = static_cast<const __nv_bfloat16*>(a);
  __nv_bfloat16* dab = static_cast<__nv_bfloat16*>(da);
  const unsigned np = (unsigned)npix, tot = (unsigned)(npix * B);
  switch (ncls) {
    case 1: head_bwd_kernel<1><<<grid, 256, 0, st>>>(ab, dy, w, np, tot, dab, dw, db); break;
    case 2: head_bwd_kernel<2><<<grid, 256, 0, st>>>(ab, dy, w, np, tot, dab, dw, db); break;
    case 3: head_bwd_kernel<3><<<grid, 256, 0, st>>>(ab, dy, w, np, tot, dab, dw, db); break;
    default: head_bwd_kernel<4><<<grid, 256, 0, st>>>(ab, dy, w, np, tot, dab, dw, db); break;
  }
  GSD_CUDA(cudaGetLastError());
  return 0;
}

static int reduce_grid(int C, long npix, int* grid, int* block) {
  const int C8 = C / 8;
  *block = 256;
  // total threads must be a multiple of C8 so that a thread always sees the same channel chunk
  long want = 148L * 2 * 256;      // 124 registers -> 2 resident blocks per SM: exactly one wave (measured best of 2..8)
  long threads = (want / C8) * C8;
  if (threads < C8) threads = C8;
  // keep it a multiple of 256 as well: lcm(C8, 256); C8 is a power of two times {1} for C in {64..1024}
  long step = C8 > 256 ? C8 : 256;
  threads = (threads / step) * step;
  if (threads < step) threads = step;
  long maxthreads = ((npix * C8 + step - 1) / step) * step;
  if (threads > maxthreads) threads = maxthreads;
  *grid = (int)(threads / 256);
  if (*grid < 1) *grid = 1;
  return 0;
}

// sums[0..C) = sum g, sums[C..2C) = sum g*zhat (accumulated; caller zeroes).  a/z/mean/rstd may be NULL together:
// then it is a plain per-channel sum of `da`.
extern "C" int gsd_op_bn_bwd_reduce(const void* da, const float* scale, const float* shift, const void* z, const float* mean,
                                    const float* rstd, long long npix, int C, float* sums, void* stream) {
  GSD_CHECK(da && sums && C % 8 == 0 && ((C / 8) & (C / 8 - 1)) == 0, "gsd_op_bn_bwd_reduce: C/8 must be a power of two");
  GSD_DEVICE_OF(da);
  int grid, block;
  reduce_grid(C, (long)npix, &grid, &block);
  bn_bwd_reduce_kernel<<<grid, block, 2 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(da), scale, shift, static_cast<const __nv_bfloat16*>(z), mean, rstd, (long)npix, C, C, sums);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_bn_bwd_apply(const void* da, const float* scale, const float* shift, const void* z, const float* mean,
                                   const float* rstd, const float* gamma, const float* sums, double count, long long npix, int C,
                                   void* dz, void* stream) {
  GSD_CHECK(da && scale && shift && z && mean && rstd && gamma && sums && dz && C % 8 == 0, "gsd_op_bn_bwd_apply: bad argument");
  GSD_DEVICE_OF(da);
  GSD_CHECK(((C / 8) & (C / 8 - 1)) == 0, "gsd_op_bn_bwd_apply: C/8 must be a power of two");
  int grid, block;
  reduce_grid(C, (long)npix, &grid, &block);
  bn_bwd_apply_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(da), scale, shift, static_cast<const __nv_bfloat16*>(z), mean, rstd, gamma, sums, (float)count,
      (long)npix, C, C, static_cast<__nv_bfloat16*>(dz));
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_maxpool_bwd(const void* a, const void* dpool, const void* dskip, int B, int H, int W, int C, void* dfull,
                                  void* stream) {
  GSD_CHECK(a && dpool && dfull && C % 8 == 0, "gsd_op_maxpool_bwd: bad argument");
  GSD_DEVICE_OF(a);
  const int C8 = C / 8;
  GSD_CHECK((C8 & (C8 - 1)) == 0 && C8 <= 256, "gsd_op_maxpool_bwd: C/8 must be a power of two <= 256");
  int c8_shift = 0;
  while ((1 << c8_shift) < C8) ++c8_shift;
  const long rows = (long)B * ((H + 1) / 2);
  const int grid = (int)(rows < 148 * 24 ? rows : 148 * 24);     // measured best of 3..32
  maxpool_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(dpool), static_cast<const __nv_bfloat16*>(dskip), C, B, H, W, C,
      c8_shift, static_cast<__nv_bfloat16*>(dfull));
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// mode 0: Conv2d (O,I,3,3) -> forward operand [O][9][Ipad];  1: -> dgrad operand [I][9][O] (flipped taps)
// mode 2: ConvTranspose2d (I,O,2,2) -> forward operand [(g)*O+o][I];  3: -> its dgrad operand [I][(g, o)]
extern "C" int gsd_op_pack_weight(int mode, const float* w, int O, int I, int Ipad, void* out, void* stream) {
  GSD_CHECK(w && out, "gsd_op_pack_weight: null argument");
  GSD_DEVICE_OF(w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (mode) {
    case 0: pack_conv_weight_kernel<<<ew_grid((long)O * 9 * Ipad), 256, 0, st>>>(w, O, I, 9, Ipad, o); break;
    case 1: pack_dgrad_weight_kernel<<<ew_grid((long)O * 9 * I), 256, 0, st>>>(w, O, I, o); break;
    case 2: pack_convt_weight_kernel<<<ew_grid(4L * O * I), 256, 0, st>>>(w, I, O, o); break;
    case 3: pack_convt_dgrad_weight_kernel<<<ew_grid(4L * O * I), 256, 0, st>>>(w, I, O, o); break;
    default: return fail(-1, "gsd_op_pack_weight: unknown mode %d", mode);
  }
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_unpack_wgrad(float* dwk, int O, int I, int Ipad, float* grad, int clear, void* stream) {
  GSD_CHECK(dwk && grad, "gsd_op_unpack_wgrad: null argument");
  GSD_DEVICE_OF(dwk);
  unpack_wgrad_kernel<<<ew_grid((long)O * I * 9), 256, 0, static_cast<cudaStream_t>(stream)>>>(dwk, O, I, Ipad, grad, clear);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" long long gsd_pack_item_units(int mode, int O, int I, int Ipad) { return pack_item_units(mode, O, I, Ipad); }

extern "C" int gsd_op_pack_weights_batched(const gsd_pack_item* items_dev, int n_items, long long total_units, void* stream) {
  GSD_CHECK(items_dev && n_items > 0 && n_items <= 64 && total_units > 0, "gsd_op_pack_weights_batched: need 1..64 items");
  GSD_DEVICE_OF(items_dev);
  static_assert(sizeof(gsd_pack_item) == sizeof(PackItemDev), "gsd_pack_item layout");
  const int grid = (int)(total_units < 148 * 8 ? total_units : 148 * 8);
  pack_weights_batched_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const PackItemDev*>(items_dev),
                                                                                 n_items, total_units);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// 1 - beta the way torch.optim.Adam evaluates it: the caller's beta is a double-precision Python float (0.999) that reached us
// as its nearest float; round it back to the shortest decimal (7 significant digits) before subtracting in double
static float one_minus(float beta) {
  char buf[32];
  snprintf(buf, sizeof buf, "%.7g", (double)beta);
  return (float)(1.0 - atof(buf));
}

// Adam (coupled L2) + EMA over a flat fp32 arena of n elements; step is 1-based; shadow may be NULL (no EMA).
// grad_scale multiplies the gradient first (1/world_size after a sum all-reduce).
// Graph-replayable variant: `counter` = 2 device int64 (Adam steps and EMA updates done so far); the kernel derives
// the bias corrections / EMA warm-up from it and a second tiny kernel advances it.
extern "C" int gsd_op_adam_ema_dev(float* p, const float* g, float* m, float* v, float* shadow, long long n, float lr, float beta1,
                                   float beta2, float eps, float weight_decay, float ema_decay, long long* counter, float grad_scale,
                                   void* stream) {
  GSD_CHECK(p && g && m && v && counter && n > 0, "gsd_op_adam_ema_dev: bad argument");
  GSD_DEVICE_OF(p);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  adam_ema_kernel<<<ew_grid((long)n / 4 + 1), 256, 0, st>>>(p, g, m, v, shadow, (long)n, lr, beta1, beta2, eps, weight_decay, 1.f, 1.f, 0.f,
                                                            grad_scale, counter, ema_decay, one_minus(beta1), one_minus(beta2));
  adam_tick_kernel<<<1, 1, 0, st>>>(counter);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_op_adam_ema(float* p, const float* g, float* m, float* v, float* shadow, long long n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, long long step, float ema_decay, long long ema_updates,
                               float grad_scale, void* stream) {
  GSD_CHECK(p && g && m && v && n > 0 && step >= 1, "gsd_op_adam_ema: bad argument");
  GSD_DEVICE_OF(p);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  double d = ema_decay;
  const double warm = (1.0 + (double)ema_updates) / (10.0 + (double)ema_updates);   // torch_ema use_num_updates
  if (warm < d) d = warm;
  adam_ema_kernel<<<ew_grid((long)n / 4 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, shadow, (long)n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), (float)(1.0 - d), grad_scale, nullptr,
      ema_decay, one_minus(beta1), one_minus(beta2));
  GSD_CUDA(cudaGetLastError());
  return 0;
}
