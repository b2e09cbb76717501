// fp32 parity variant of the training plan (included by train_plan.h; kernels in train_f32.cuh).  Same step, same
// parameter / gradient / bucket protocol as the bf16 plan, every launch on the caller's stream:
//   forward : operand packing, NCHW -> NHWC, per unit conv (FFMA implicit GEMM) -> batch statistics (double) -> finalize ->
//             BatchNorm + ReLU (-> max-pool), transposed convs with bias, OutConv
//   backward: OutConv, per unit BatchNorm/ReLU backward (reduce in double, apply) -> input gradient (conv with the
//             transposed, flipped filter) -> weight gradient (pixel-split GEMM, fp32 atomics), transposed-conv backward,
//             max-pool backward with the skip-connection gradient added
// Reference: train_utils/train_unet.py:346-377 (loop body), train_utils/unet.py:7-88 (modules).
#pragma once
#include "train_f32.cuh"

namespace {

inline int f32_grid(long n) {
  long b = (n + 255) / 256;
  if (b > 148L * 16) b = 148L * 16;
  return (int)(b < 1 ? 1 : b);
}

#define F32_LAUNCHED(p)          \
  do {                           \
    GSD_CUDA(cudaGetLastError()); \
    ++(p)->launches_f32;         \
  } while (0)

// out = conv(src) over `ncols` output channels starting at column `col0` of w ([ntaps][Cin][ldw])
int f32_conv(gsd_train_plan* p, const float* src0, int C0, const float* src1, int C1, int H1, int W1, int off_y, int off_x, int B, int H, int W,
             const float* w, int ldw, int col0, int ncols, int ntaps, int groups, const float* shift, float* out, cudaStream_t st,
             int in_stride = 0, int sH = 0, int sW = 0, int s_oy = 0, int s_ox = 0) {
  F32Conv c;
  c.src0 = src0; c.C0 = C0; c.src1 = src1; c.C1 = C1; c.H1 = H1; c.W1 = W1; c.off_y = off_y; c.off_x = off_x;
  c.w = w + col0; c.scale = nullptr; c.shift = shift; c.out = out;
  c.B = B; c.H = H; c.W = W; c.Cout = ncols; c.groups = groups; c.ntaps = ntaps; c.relu = 0;
  c.ldw = ldw; c.in_stride = in_stride; c.sH = sH; c.sW = sW; c.s_oy = s_oy; c.s_ox = s_ox;
  for (int t = 0; t < 9; ++t) { c.dy[t] = 0; c.dx[t] = 0; }
  if (ntaps == 9) {
    for (int t = 0; t < 9; ++t) { c.dy[t] = (int8_t)(t / 3 - 1); c.dx[t] = (int8_t)(t % 3 - 1); }
  } else if (ntaps == 4) {
    for (int t = 0; t < 4; ++t) { c.dy[t] = (int8_t)(t >> 1); c.dx[t] = (int8_t)(t & 1); }
  }
  const long M = (long)B * H * W;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((groups * ncols + 63) / 64));
  conv_f32_kernel<<<grid, 256, 0, st>>>(c);
  F32_LAUNCHED(p);
  return 0;
}

int f32_chan_stats(gsd_train_plan* p, const float* t, int B, int H, int W, int C, int oy, int ox, int h2, int w2, int want_sq, double* out,
                   cudaStream_t st) {
  GSD_CHECK(C <= 1024, "fp32 training path: more than 1024 channels");
  long blocks = ((long)B * h2 * w2 + 511) / 512;
  if (blocks > 148 * 8) blocks = 148 * 8;
  chan_stats_f32_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(t, B, H, W, C, oy, ox, h2, w2, want_sq, out);
  F32_LAUNCHED(p);
  return 0;
}

int f32_wgrad(gsd_train_plan* p, F32Wgrad& g, cudaStream_t st) {
  const long M = (long)g.B * g.H * g.W;
  g.chunk = 2048;
  const int Ctot = g.C0 + g.C1;
  dim3 grid((unsigned)(((Ctot + 63) / 64) * g.ntaps), (unsigned)((g.Cout + 63) / 64), (unsigned)((M + g.chunk - 1) / g.chunk));
  GSD_CHECK(grid.z <= 65535, "fp32 training path: too many pixel chunks");
  wgrad_f32_kernel<<<grid, 256, 0, st>>>(g);
  F32_LAUNCHED(p);
  return 0;
}

int f32_unit_forward(gsd_train_plan* p, TrainUnit& u, cudaStream_t st) {
  const int B = p->g.batch;
  const long n = (long)B * u.H * u.W * u.Cout;
  float* z = wsp<float>(p, u.z);
  float* consts = wsp<float>(p, u.consts);
  GSD_TRY(f32_conv(p, wsp<float>(p, u.src0), u.C0, u.C1 ? wsp<float>(p, u.src1) : nullptr, u.C1, u.H1, u.W1, u.off_y, u.off_x, B, u.H, u.W,
                   wsp<float>(p, u.w_fwd), u.Cout, 0, u.Cout, 9, 1, nullptr, z, st));
  GSD_TRY(f32_chan_stats(p, z, B, u.H, u.W, u.Cout, 0, 0, u.H, u.W, 1, wsp<double>(p, u.dstat), st));
  bn_finalize_f32_kernel<<<(u.Cout + 127) / 128, 128, 0, st>>>(wsp<double>(p, u.dstat), (double)B * u.H * u.W, p->params[u.p_g], p->params[u.p_b],
                                                               p->bnbuf[2 * u.bn], p->bnbuf[2 * u.bn + 1], 0.1f, 1e-5f, u.Cout, consts,
                                                               p->nbt.empty() ? nullptr : p->nbt[u.bn]);
  F32_LAUNCHED(p);
  bn_relu_apply_f32_kernel<<<f32_grid(n), 256, 0, st>>>(z, consts, n, u.Cout, wsp<float>(p, u.a));
  F32_LAUNCHED(p);
  if (u.pool) {
    maxpool_f32_kernel<<<f32_grid((long)B * (u.H / 2) * (u.W / 2) * u.Cout), 256, 0, st>>>(wsp<float>(p, u.a), B, u.H, u.W, u.Cout,
                                                                                           wsp<float>(p, u.pooled));
    F32_LAUNCHED(p);
  }
  return 0;
}

int f32_forward(gsd_train_plan* p, const float* x, float* y, cudaStream_t st) {
  const gsd_geometry& g = p->g;
  const int B = g.batch;
  p->launches_f32 = 0;
  // operands: forward [tap][Cin][Cout] and input-gradient [8 - tap][Cout][Cin] forms of every filter
  auto pack_unit = [&](TrainUnit& u) -> int {
    const int I = u.C0 + u.C1;
    pack_conv_weight_f32_kernel<<<f32_grid(9L * u.Cout * I), 256, 0, st>>>(p->params[u.p_w], u.Cout, I, 9, wsp<float>(p, u.w_fwd));
    F32_LAUNCHED(p);
    if (!u.first) {
      pack_conv_dgrad_f32_kernel<<<f32_grid(9L * u.Cout * I), 256, 0, st>>>(p->params[u.p_w], u.Cout, I, wsp<float>(p, u.w_dgrad));
      F32_LAUNCHED(p);
    }
    return 0;
  };
  for (auto& u : p->enc) GSD_TRY(pack_unit(u));
  for (auto& u : p->dec) GSD_TRY(pack_unit(u));
  for (auto& t : p->ups) {
    pack_convt_weight_f32_kernel<<<f32_grid(4L * t.Cin * t.Cout), 256, 0, st>>>(p->params[t.p_w], t.Cin, t.Cout, wsp<float>(p, t.w_fwd));
    F32_LAUNCHED(p);
    pack_convt_dgrad_f32_kernel<<<f32_grid(4L * t.Cin * t.Cout), 256, 0, st>>>(p->params[t.p_w], t.Cin, t.Cout, wsp<float>(p, t.w_dgrad));
    F32_LAUNCHED(p);
  }
  GSD_CUDA(cudaMemsetAsync(p->ws + p->zero_arena, 0, p->zero_bytes, st));
  GSD_CUDA(cudaMemsetAsync(p->ws + p->dzero, 0, p->dzero_bytes, st));
  if (p->n_prep_fwd) {
    prep_kernel<<<p->n_prep_fwd, 256, 0, st>>>(wsp<PrepItem>(p, p->prep_fwd));     // transposed-conv bias for each of the 4 (dy,dx) groups
    F32_LAUNCHED(p);
  }
  const long npix = (long)g.height * g.width;
  nchw_to_nhwc_f32_kernel<<<f32_grid((long)B * g.in_channels * npix), 256, 0, st>>>(x, B, g.in_channels, npix, wsp<float>(p, p->in16));
  F32_LAUNCHED(p);
  for (auto& u : p->enc) GSD_TRY(f32_unit_forward(p, u, st));
  for (int i = 0; i < p->depth; ++i) {
    TrainUp& t = p->ups[i];
    GSD_TRY(f32_conv(p, wsp<float>(p, t.src), t.Cin, nullptr, 0, 0, 0, 0, 0, B, t.hs, t.ws, wsp<float>(p, t.w_fwd), 4 * t.Cout, 0, t.Cout, 1, 4,
                     wsp<float>(p, t.bias4), wsp<float>(p, t.u), st));
    GSD_TRY(f32_unit_forward(p, p->dec[2 * i], st));
    GSD_TRY(f32_unit_forward(p, p->dec[2 * i + 1], st));
  }
  TrainUnit& last = p->dec.back();
  head_f32_kernel<<<f32_grid(npix * B * g.n_classes), 256, 0, st>>>(wsp<float>(p, last.a), g.dims[0], p->params[p->n_params - 2],
                                                                   p->params[p->n_params - 1], g.n_classes, 1.f, 0.f, npix, B, y);
  F32_LAUNCHED(p);
  return 0;
}

// backward of conv -> BatchNorm -> ReLU; da: gradient of the unit's activation
int f32_unit_backward(gsd_train_plan* p, TrainUnit& u, const float* da, bool need_dx, cudaStream_t st) {
  const int B = p->g.batch;
  const long npix = (long)B * u.H * u.W, n = npix * u.Cout;
  const float* a = wsp<float>(p, u.a);
  const float* z = wsp<float>(p, u.z);
  const float* consts = wsp<float>(p, u.consts);
  double* sums = wsp<double>(p, u.dbwd);
  float* dz = wsp<float>(p, u.dz);
  long blocks = (npix + 511) / 512;
  if (blocks > 148 * 8) blocks = 148 * 8;
  bn_bwd_reduce_f32_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(da, a, z, consts, npix, u.Cout, sums);
  F32_LAUNCHED(p);
  bn_bwd_apply_f32_kernel<<<f32_grid(n), 256, 0, st>>>(da, a, z, consts, sums, (double)npix, n, u.Cout, dz, p->grads[u.p_g], p->grads[u.p_b]);
  F32_LAUNCHED(p);
  grad_done(p, u.p_b);
  grad_done(p, u.p_g);
  const int Ctot = u.C0 + u.C1;
  if (need_dx) {
    const float* wd = wsp<float>(p, u.w_dgrad);               // [8 - tap][Cout][Ctot]
    GSD_TRY(f32_conv(p, dz, u.Cout, nullptr, 0, 0, 0, 0, 0, B, u.H, u.W, wd, Ctot, 0, u.C0, 9, 1, nullptr, wsp<float>(p, u.din0), st));
    if (u.C1)
      GSD_TRY(f32_conv(p, dz, u.Cout, nullptr, 0, 0, 0, 0, 0, B, u.H, u.W, wd, Ctot, u.C0, u.C1, 9, 1, nullptr, wsp<float>(p, u.din1), st));
  }
  GSD_CUDA(cudaMemsetAsync(p->grads[u.p_w], 0, (size_t)u.Cout * Ctot * 9 * 4, st));
  F32Wgrad g{};
  g.a0 = wsp<float>(p, u.src0); g.C0 = u.C0; g.aH = u.H; g.aW = u.W;
  g.a1 = u.C1 ? wsp<float>(p, u.src1) : nullptr; g.C1 = u.C1; g.H1 = u.H1; g.W1 = u.W1; g.off_y = u.off_y; g.off_x = u.off_x;
  g.d = dz; g.Cout = u.Cout; g.dH = u.H; g.dW = u.W;
  g.B = B; g.H = u.H; g.W = u.W; g.ntaps = 9; g.d_stride = 1; g.d_oy = 0; g.d_ox = 0;
  for (int t = 0; t < 9; ++t) { g.a_dy[t] = (int8_t)(t / 3 - 1); g.a_dx[t] = (int8_t)(t % 3 - 1); g.d_dy[t] = 0; g.d_dx[t] = 0; }
  g.grad = p->grads[u.p_w]; g.so = (long)Ctot * 9; g.si = 9; g.st = 1;          // Conv2d weight (O, I, 3, 3)
  GSD_TRY(f32_wgrad(p, g, st));
  grad_done(p, u.p_w);
  return 0;
}

int f32_backward(gsd_train_plan* p, const float* dy, cudaStream_t st) {
  const gsd_geometry& G = p->g;
  const int B = G.batch, depth = p->depth, ncls = G.n_classes, C0 = G.dims[0];
  const long npix_img = (long)G.height * G.width;
  // every double accumulator of the backward pass starts from zero, also when backward runs twice on one forward (the
  // forward's batch-statistics sums in the same arena are dead by now)
  GSD_CUDA(cudaMemsetAsync(p->ws + p->dzero, 0, p->dzero_bytes, st));
  // OutConv (unet.py:54-57)
  TrainUnit& last = p->dec.back();
  float* da_head = wsp<float>(p, p->da_head);
  double* dh = wsp<double>(p, p->dhead);
  {
    long blocks = (npix_img * B + 511) / 512;
    if (blocks > 148 * 8) blocks = 148 * 8;
    head_bwd_f32_kernel<<<(int)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(wsp<float>(p, last.a), dy, p->params[p->n_params - 2], ncls, C0, npix_img, B,
                                                                        da_head, dh);
    F32_LAUNCHED(p);
    f64_to_f32_kernel<<<(ncls * C0 + 255) / 256, 256, 0, st>>>(dh, ncls * C0, p->grads[p->n_params - 2]);
    F32_LAUNCHED(p);
    f64_to_f32_kernel<<<1, 256, 0, st>>>(dh + ncls * C0, ncls, p->grads[p->n_params - 1]);
    F32_LAUNCHED(p);
  }
  const float* da = da_head;
  for (int i = depth - 1; i >= 0; --i) {
    TrainUnit& u1 = p->dec[2 * i];
    TrainUnit& u2 = p->dec[2 * i + 1];
    GSD_TRY(f32_unit_backward(p, u2, da, true, st));
    if (i == depth - 1) { grad_done(p, p->n_params - 1); grad_done(p, p->n_params - 2); }
    GSD_TRY(f32_unit_backward(p, u1, wsp<float>(p, u2.din0), true, st));
    TrainUp& t = p->ups[i];
    const float* du = wsp<float>(p, u1.din1);          // (B, H, W, t.Cout); the transposed conv's output is the (2hs x 2ws) window at (off_y, off_x)
    GSD_TRY(f32_conv(p, du, t.Cout, nullptr, 0, 0, 0, 0, 0, B, t.hs, t.ws, wsp<float>(p, t.w_dgrad), t.Cin, 0, t.Cin, 4, 1, nullptr,
                     wsp<float>(p, t.din), st, 2, u1.H, u1.W, u1.off_y, u1.off_x));
    GSD_TRY(f32_chan_stats(p, du, B, u1.H, u1.W, t.Cout, u1.off_y, u1.off_x, 2 * t.hs, 2 * t.ws, 0, wsp<double>(p, t.dbias), st));
    f64_to_f32_kernel<<<(t.Cout + 255) / 256, 256, 0, st>>>(wsp<double>(p, t.dbias), t.Cout, p->grads[t.p_b]);
    F32_LAUNCHED(p);
    grad_done(p, t.p_b);
    GSD_CUDA(cudaMemsetAsync(p->grads[t.p_w], 0, (size_t)4 * t.Cin * t.Cout * 4, st));
    F32Wgrad g{};
    g.a0 = wsp<float>(p, t.src); g.C0 = t.Cin; g.aH = t.hs; g.aW = t.ws; g.a1 = nullptr; g.C1 = 0;
    g.d = du; g.Cout = t.Cout; g.dH = u1.H; g.dW = u1.W;
    g.B = B; g.H = t.hs; g.W = t.ws; g.ntaps = 4; g.d_stride = 2; g.d_oy = u1.off_y; g.d_ox = u1.off_x;
    for (int k = 0; k < 4; ++k) { g.a_dy[k] = 0; g.a_dx[k] = 0; g.d_dy[k] = (int8_t)(k >> 1); g.d_dx[k] = (int8_t)(k & 1); }
    g.grad = p->grads[t.p_w]; g.so = 4; g.si = (long)t.Cout * 4; g.st = 1;       // ConvTranspose2d weight (I, O, 2, 2)
    GSD_TRY(f32_wgrad(p, g, st));
    grad_done(p, t.p_w);
    da = wsp<float>(p, t.din);
  }
  for (int l = depth; l >= 0; --l) {
    TrainUnit& u1 = p->enc[2 * l];
    TrainUnit& u2 = p->enc[2 * l + 1];
    if (l < depth) {
      float* dfull = wsp<float>(p, p->dfull[l]);
      maxpool_bwd_f32_kernel<<<f32_grid((long)B * u2.H * u2.W * u2.Cout), 256, 0, st>>>(wsp<float>(p, u2.a), wsp<float>(p, p->enc[2 * l + 2].din0),
                                                                                        wsp<float>(p, p->dec[2 * (depth - 1 - l)].din0), B, u2.H,
                                                                                        u2.W, u2.Cout, dfull);
      F32_LAUNCHED(p);
      da = dfull;
    }
    GSD_TRY(f32_unit_backward(p, u2, da, true, st));
    GSD_TRY(f32_unit_backward(p, u1, wsp<float>(p, u2.din0), l > 0, st));
  }
  return 0;
}

}  // namespace
