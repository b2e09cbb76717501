// Plan-level training ABI (include/gsd_b200.h "training plan"): the loop body of train_utils/train_unet.py:346-377
//     output = unet(x=input); loss = MSE(output, target); loss.backward(); optimizer.step(); ema.update()
// sequenced in C++ over caller-owned memory.  One gsd_train_plan per (geometry, device): it lays out every saved
// activation, gradient temporary, packed operand and per-layer constant inside ONE caller-allocated workspace, and
// gsd_train_forward / gsd_backward / gsd_adam_ema_step / gsd_train_step enqueue the whole step as library kernels (plus a
// handful of memset nodes) on the caller's stream -- no allocation, no host synchronisation, no framework kernels, so
// the step can be captured into a CUDA graph as is.  Weight-gradient GEMMs run on a plan-owned side stream (forked /
// joined with events, which capture follows), concurrently with the next unit's HBM-bound BatchNorm-backward passes.
// Data parallelism: the caller supplies a bucket id per parameter; the plan counts gradients down in backward order and
// calls back (on the host, at enqueue time) when the last gradient kernel of a bucket has been launched, so the caller can
// order its communication stream after both compute streams and enqueue the NCCL all-reduce of that arena range.
#pragma once
#include <vector>

#include "train_abi.h"

namespace {

struct TrainUnit {          // conv3x3 -> BatchNorm2d -> ReLU (unet.py:11-13 / 14-16)
  int H = 0, W = 0, C0 = 0, C1 = 0, Cout = 0, cin_real = 0;
  int H1 = 0, W1 = 0, off_y = 0, off_x = 0;
  bool first = false, pool = false, apply = true;
  bool one_dgrad = false;      // concat unit: both halves of the input gradient from ONE launch (split-output halo kernel)
  int p_w = 0, p_g = 0, p_b = 0, bn = 0, neg_off = 0;
  size_t z = 0, a = 0, pooled = 0, consts = 0, stats = 0, dz = 0, din0 = 0, din1 = 0, w_fwd = 0, w_dgrad = 0, dwk = 0;
  size_t src0 = 0, src1 = 0;      // workspace offsets of the conv inputs
  size_t dstat = 0, dbwd = 0;     // fp32 parity path: double [2C] accumulators (batch statistics / their gradients)
};

struct TrainUp {            // ConvTranspose2d(k=2, s=2) of one decoder block (unet.py:36)
  int Cin = 0, Cout = 0, hs = 0, ws = 0, p_w = 0, p_b = 0;
  size_t u = 0, w_fwd = 0, w_dgrad = 0, bias4 = 0, din = 0, src = 0;
  size_t dbias = 0;               // fp32 parity path: double [Cout]
};

struct PrepItem {           // dst[i] = src ? mul * src[i % csrc] : 0   (per-step constants and small zero fills, one launch)
  const float* src;
  float* dst;
  int n, csrc;
  float mul;
};

__global__ void __launch_bounds__(256) prep_kernel(const PrepItem* __restrict__ items) {
  const PrepItem it = items[blockIdx.x];
  for (int i = threadIdx.x; i < it.n; i += blockDim.x) it.dst[i] = it.src ? it.mul * it.src[i % it.csrc] : 0.f;
}

// zero the F.pad frame (unet.py:43-47) of a dense (B, Hf, Wf, C) bf16 gradient whose (h2 x w2) window at (oy, ox) is the
// transposed conv's output gradient: the bias gradient then is a plain per-channel sum of the tensor
__global__ void __launch_bounds__(256) zero_frame_kernel(__nv_bfloat16* __restrict__ t, int B, int Hf, int Wf, int C8, int oy, int ox,
                                                         int h2, int w2) {
  const int top = oy * Wf, bottom = (Hf - oy - h2) * Wf, left = h2 * ox, right = h2 * (Wf - ox - w2);
  const int per_img = top + bottom + left + right;
  const long total = (long)B * per_img * C8;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % C8);
    long r = idx / C8;
    const int b = (int)(r / per_img);
    int k = (int)(r - (long)b * per_img), y, x;
    if (k < top) { y = k / Wf; x = k - y * Wf; }
    else if ((k -= top) < bottom) { y = oy + h2 + k / Wf; x = k % Wf; }
    else if ((k -= bottom) < left) { y = oy + k / ox; x = k % ox; }
    else { k -= left; const int wr = Wf - ox - w2; y = oy + k / wr; x = ox + w2 + k % wr; }
    *reinterpret_cast<uint4*>(t + (((long)b * Hf + y) * Wf + x) * (C8 * 8L) + c8 * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace

static bool& dry_create_mode() {
  static thread_local bool on = false;
  return on;
}

struct gsd_train_plan {
  gsd_geometry g{};
  int device = 0, num_sms = 148, depth = 0;
  std::vector<int> Hs, Ws;
  std::vector<TrainUnit> enc, dec;      // 2 per level / decoder block
  std::vector<TrainUp> ups;
  int n_params = 0, n_bn = 0, neg_total = 0;
  std::vector<long long> param_numel;
  // workspace layout (byte offsets)
  size_t in16 = 0, y = 0, dy = 0, zero_arena = 0, zero_bytes = 0, loss = 0, neg = 0, packed = 0, packed_bytes = 0, dwk0 = 0, dwk_bytes = 0,
         pack_table = 0, prep_fwd = 0, prep_bwd = 0, ws_bytes = 0;
  std::vector<size_t> dfull;            // max-pool backward outputs per level
  int n_pack = 0, n_prep_fwd = 0, n_prep_bwd = 0;
  long long pack_units = 0;
  // bound state
  char* ws = nullptr;
  std::vector<const float*> params;
  std::vector<float*> grads;
  std::vector<float*> bnbuf;
  std::vector<long long*> nbt;
  bool bound = false;
  // buckets
  std::vector<int> bucket_of;
  std::vector<long long> bucket_lo, bucket_hi;
  std::vector<int> bucket_size, pending;
  // streams
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> events;
  size_t ev_next = 0;
  bool forked = false;
  int overlap = 1;
  bool f32 = false;      // geometry.dtype == GSD_DTYPE_FP32: the FFMA parity path (train_plan_f32.h), everything on the main stream
  size_t dzero = 0, dzero_bytes = 0, dhead = 0, da_head = 0;   // fp32: double accumulator arena, OutConv gradient accumulators / input gradient
  int launches_f32 = 0;  // launches of the last fp32 step (counted while enqueuing)
  bool dry = false;      // gsd_debug_train_plan_create: walk the step's structure (gradient order, bucket callbacks) without a GPU
  // per-call
  gsd_bucket_cb cb = nullptr;
  void* cb_user = nullptr;
  cudaStream_t main = nullptr;
};

// every launch of the backward pass goes through this: the dry-run plan of the CPU tests enqueues nothing
#define TP_RUN(p, expr)            \
  do {                             \
    if (!(p)->dry) GSD_TRY(expr);  \
  } while (0)

namespace {

size_t tbump(size_t* cur, size_t bytes) {
  const size_t off = align_up(*cur, 1024);
  *cur = off + bytes;
  return off;
}

template <class T>
T* wsp(const gsd_train_plan* p, size_t off) { return reinterpret_cast<T*>(p->ws + off); }

cudaEvent_t next_event(gsd_train_plan* p) {
  cudaEvent_t e = p->events[p->ev_next];
  p->ev_next = (p->ev_next + 1) % p->events.size();
  return e;
}

// kernels launched between fork() and the matching unfork() go to the side stream, ordered after everything issued so
// far on the main stream
int side_begin(gsd_train_plan* p, cudaStream_t* st) {
  if (!p->overlap) { *st = p->main; return 0; }
  if (p->dry) { p->forked = true; *st = p->side; return 0; }
  cudaEvent_t e = next_event(p);
  GSD_CUDA(cudaEventRecord(e, p->main));
  GSD_CUDA(cudaStreamWaitEvent(p->side, e, 0));
  p->forked = true;
  *st = p->side;
  return 0;
}
int side_join(gsd_train_plan* p) {
  if (!p->forked) return 0;
  if (p->dry) { p->forked = false; return 0; }
  cudaEvent_t e = next_event(p);
  GSD_CUDA(cudaEventRecord(e, p->side));
  GSD_CUDA(cudaStreamWaitEvent(p->main, e, 0));
  p->forked = false;
  return 0;
}

void grad_done(gsd_train_plan* p, int pi) {
  if (p->bucket_of.empty()) return;
  const int b = p->bucket_of[pi];
  if (--p->pending[b] == 0 && p->cb) p->cb(p->cb_user, b, p->bucket_lo[b], p->bucket_hi[b], p->main, p->forked ? p->side : nullptr);
}

}  // namespace
#include "train_plan_f32.h"
namespace {

int bn_bwd_launch(const void* da, const float* scale, const float* shift, const void* z, const float* mean, const float* rstd,
                  const float* gamma, double count, long npix, int C, float* dbeta, float* dgamma, void* dz, cudaStream_t st) {
  int grid, block;
  reduce_grid(C, npix, &grid, &block);
  bn_bwd_reduce_kernel<<<grid, block, 2 * C * sizeof(float), st>>>(static_cast<const __nv_bfloat16*>(da), scale, shift,
                                                                   static_cast<const __nv_bfloat16*>(z), mean, rstd, npix, C, C, dbeta, dgamma);
  bn_bwd_apply_kernel<<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(da), scale, shift, static_cast<const __nv_bfloat16*>(z), mean,
                                              rstd, gamma, dbeta, (float)count, npix, C, C, static_cast<__nv_bfloat16*>(dz), dgamma);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

int unit_forward(gsd_train_plan* p, TrainUnit& u, cudaStream_t st) {
  const int B = p->g.batch;
  float* consts = wsp<float>(p, u.consts);
  float* scale = consts, *shift = consts + u.Cout, *mean = consts + 2 * u.Cout, *rstd = consts + 3 * u.Cout;
  float* stats = wsp<float>(p, u.stats);
  const float* neg = wsp<float>(p, p->neg) + u.neg_off;
  // z is stored centred on the running mean (bf16 then rounds relative to the fluctuation of z, not to its mean); the batch
  // statistics are taken from the raw fp32 accumulators in the conv epilogue
  GSD_TRY(gsd_op_conv_auto_bf16(p->ws + u.src0, u.C0, u.C1 ? p->ws + u.src1 : nullptr, u.C1, u.H1, u.W1, u.off_y, u.off_x, B, u.H, u.W,
                                p->ws + u.w_fwd, u.Cout, 9, 1, nullptr, neg, 0, p->ws + u.z, nullptr, stats, p->device, st));
  GSD_TRY(gsd_op_bn_finalize(stats, (double)B * u.H * u.W, p->params[u.p_g], p->params[u.p_b], p->bnbuf[2 * u.bn], p->bnbuf[2 * u.bn + 1], 0.1f,
                             1e-5f, u.Cout, neg, scale, shift, mean, rstd, p->nbt.empty() ? nullptr : p->nbt[u.bn], st));
  if (u.apply)
    GSD_TRY(gsd_op_bn_relu_apply(p->ws + u.z, scale, shift, B, u.H, u.W, u.Cout, p->ws + u.a, u.pool ? p->ws + u.pooled : nullptr, st));
  return 0;
}

// backward of conv -> BN -> ReLU; `da` = gradient of the unit's output (null for the last unit: the OutConv backward is
// fused in and reads `dy`).  Input gradients go to u.din0 (and u.din1 for the concat's second source).
int unit_backward(gsd_train_plan* p, TrainUnit& u, const void* da, const float* dy, bool need_dx) {
  const int B = p->g.batch;
  cudaStream_t st = p->main;
  float* consts = wsp<float>(p, u.consts);
  float* scale = consts, *shift = consts + u.Cout, *mean = consts + 2 * u.Cout, *rstd = consts + 3 * u.Cout;
  float* dgamma = p->grads[u.p_g], *dbeta = p->grads[u.p_b];
  const long npix = (long)B * u.H * u.W;
  if (!da) {
    const int pw = p->n_params - 2, pb = p->n_params - 1;     // outc.conv.weight / bias
    const float* w_head = p->params[pw];
    const unsigned np = (unsigned)(u.H * u.W), tot = (unsigned)npix;
    const __nv_bfloat16* zb = wsp<__nv_bfloat16>(p, u.z);
    __nv_bfloat16* dzb = wsp<__nv_bfloat16>(p, u.dz);
    float* dw = p->grads[pw], *db = p->grads[pb];
    switch (p->g.n_classes) {
      case 1: TP_RUN(p, launch_head_bn_bwd<1>(zb, dy, w_head, scale, shift, mean, rstd, p->params[u.p_g], (float)npix, np, tot, dbeta, dw, db, dzb, st, dgamma)); break;
      case 2: TP_RUN(p, launch_head_bn_bwd<2>(zb, dy, w_head, scale, shift, mean, rstd, p->params[u.p_g], (float)npix, np, tot, dbeta, dw, db, dzb, st, dgamma)); break;
      case 3: TP_RUN(p, launch_head_bn_bwd<3>(zb, dy, w_head, scale, shift, mean, rstd, p->params[u.p_g], (float)npix, np, tot, dbeta, dw, db, dzb, st, dgamma)); break;
      default: TP_RUN(p, launch_head_bn_bwd<4>(zb, dy, w_head, scale, shift, mean, rstd, p->params[u.p_g], (float)npix, np, tot, dbeta, dw, db, dzb, st, dgamma)); break;
    }
  } else {
    TP_RUN(p, bn_bwd_launch(da, scale, shift, p->ws + u.z, mean, rstd, p->params[u.p_g], (double)npix, npix, u.Cout, dbeta, dgamma,
                          p->ws + u.dz, st));
  }
  grad_done(p, u.p_b);
  grad_done(p, u.p_g);
  // dgrad first (the critical path waits for it), then the weight gradient on the side stream, where it overlaps the
  // next unit's HBM-bound BatchNorm-backward passes
  if (need_dx) {
    const char* wd = p->ws + u.w_dgrad;                        // [ci][9][co] bf16
    if (u.C1) {
      // the decoder's concat conv: ONE launch with N = C0 + C1 whose epilogue writes the skip half and the up half of the
      // input gradient to two dense tensors (at the top level two N = 64 launches are bound by the smem operand reads, one
      // N = 128 launch is not: 24.35 -> 24.31 ms per step); the tap-streaming kernel and small batches keep two launches
      ConvDesc d;
      d.src0 = p->ws + u.dz; d.C0 = u.Cout; d.B = B; d.H = u.H; d.W = u.W; d.w = wd; d.Cout = u.C0 + u.C1; d.groups = 1;
      taps3x3(&d);
      d.out = p->ws + u.din0; d.out2 = p->ws + u.din1; d.split_c = u.C0;
      if (u.one_dgrad && !p->dry) {
        HaloLaunch L;
        GSD_TRY(build_halo_launch(d, p->num_sms, &L));
        GSD_TRY(run_halo_launch(L, st));
      } else if (!u.one_dgrad) {
        const size_t rows = (size_t)u.C0 * 9 * u.Cout * 2;       // bytes of the skip half
        TP_RUN(p, gsd_op_conv_auto_bf16(p->ws + u.dz, u.Cout, nullptr, 0, 0, 0, 0, 0, B, u.H, u.W, wd, u.C0, 9, 1, nullptr, nullptr, 0,
                                      p->ws + u.din0, nullptr, nullptr, p->device, st));
        TP_RUN(p, gsd_op_conv_auto_bf16(p->ws + u.dz, u.Cout, nullptr, 0, 0, 0, 0, 0, B, u.H, u.W, wd + rows, u.C1, 9, 1, nullptr, nullptr, 0,
                                      p->ws + u.din1, nullptr, nullptr, p->device, st));
      }
    } else {
      TP_RUN(p, gsd_op_conv_auto_bf16(p->ws + u.dz, u.Cout, nullptr, 0, 0, 0, 0, 0, B, u.H, u.W, wd, u.C0, 9, 1, nullptr, nullptr, 0,
                                    p->ws + u.din0, nullptr, nullptr, p->device, st));
    }
  }
  cudaStream_t sw;
  GSD_TRY(side_begin(p, &sw));
  float* dwk = wsp<float>(p, u.dwk);
  TP_RUN(p, gsd_op_wgrad3x3_bf16(p->ws + u.src0, u.C0, u.C1 ? p->ws + u.src1 : nullptr, u.C1, u.H1, u.W1, u.off_y, u.off_x, p->ws + u.dz, u.Cout,
                               B, u.H, u.W, dwk, p->device, sw));
  TP_RUN(p, gsd_op_unpack_wgrad(dwk, u.Cout, u.first ? u.cin_real : u.C0 + u.C1, u.C0 + u.C1, p->grads[u.p_w], 1, sw));
  grad_done(p, u.p_w);
  return 0;
}

}  // namespace

extern "C" int gsd_train_plan_create(gsd_train_plan** out, const gsd_geometry* g, int device) {
  GSD_CHECK(out && g, "gsd_train_plan_create: null argument");
  GSD_CHECK(g->dtype == GSD_DTYPE_BF16 || g->dtype == GSD_DTYPE_FP32, "gsd_train_plan_create: geometry.dtype must be GSD_DTYPE_BF16 or GSD_DTYPE_FP32");
  GSD_CHECK(g->mode == GSD_MODE_TRAIN, "gsd_train_plan_create: geometry.mode must be GSD_MODE_TRAIN");
  GSD_CHECK(g->batch >= 1 && g->height >= 1 && g->width >= 1, "gsd_train_plan_create: bad shape");
  GSD_CHECK(g->in_channels >= 1 && g->in_channels <= 8, "gsd_train_plan_create: in_channels %d not in 1..8", g->in_channels);
  GSD_CHECK(g->n_classes >= 1 && g->n_classes <= 4, "gsd_train_plan_create: n_classes %d not in 1..4", g->n_classes);
  GSD_CHECK(g->n_dims >= 2 && g->n_dims <= GSD_MAX_DIMS, "gsd_train_plan_create: n_dims %d not in 2..%d", g->n_dims, GSD_MAX_DIMS);
  GSD_CHECK(g->dims[0] == 64, "gsd_train_plan_create: layer_dimensions[0] must be 64");
  for (int i = 0; i + 1 < g->n_dims; ++i)
    GSD_CHECK(g->dims[i + 1] == 2 * g->dims[i], "gsd_train_plan_create: layer_dimensions must double at every level (unet.py:48)");
  const bool dry = dry_create_mode();
  int sms = 148;
  if (!dry) {
    int ndev = 0, major = 0;
    GSD_CUDA(cudaGetDeviceCount(&ndev));
    GSD_CHECK(device >= 0 && device < ndev, "gsd_train_plan_create: device %d out of range (%d devices)", device, ndev);
    GSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    GSD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    GSD_CHECK(major == 10, "gsd_train_plan_create: device %d is sm_%d0; this library contains sm_100a code only (no fallback)", device, major);
  }
  DeviceGuard guard(dry ? -1 : device);
  GSD_TRY(guard.status());

  gsd_train_plan* p = new gsd_train_plan();
  p->dry = dry;
  p->f32 = g->dtype == GSD_DTYPE_FP32;
  p->g = *g;
  p->device = device;
  p->num_sms = sms;
  p->depth = g->n_dims - 1;
  p->Hs.push_back(g->height);
  p->Ws.push_back(g->width);
  for (int l = 0; l < p->depth; ++l) { p->Hs.push_back(p->Hs.back() / 2); p->Ws.push_back(p->Ws.back() / 2); }
  if (p->Hs.back() < 1 || p->Ws.back() < 1) {
    delete p;
    return fail(-1, "gsd_train_plan_create: %dx%d is too small for %d poolings", g->height, g->width, g->n_dims - 1);
  }
  const size_t B = g->batch;
  const int* dims = g->dims;
  size_t cur = 0;
  const bool f32 = p->f32;
  const size_t es = f32 ? 4 : 2;      // activation / packed operand element size
  size_t dz_cur = 0;                  // fp32: double accumulators
  p->in16 = tbump(&cur, f32 ? B * g->height * g->width * g->in_channels * 4 : B * g->height * g->width * 16 * 2);
  const size_t out_bytes = B * g->n_classes * g->height * g->width * 4;
  p->y = tbump(&cur, out_bytes);
  p->dy = tbump(&cur, out_bytes);
  int pi = 0, bn = 0, neg = 0;
  size_t zero_cur = 0, packed_cur = 0, dwk_cur = 0;
  auto make_unit = [&](int l, int C0, int C1, int cin_real, int Cout, bool first, bool pool, bool apply, bool need_dx) {
    TrainUnit u;
    u.H = p->Hs[l]; u.W = p->Ws[l]; u.C0 = C0; u.C1 = C1; u.cin_real = cin_real; u.Cout = Cout;
    u.first = first; u.pool = pool; u.apply = apply || f32;
    if (f32 && first) u.C0 = cin_real;       // no 16-channel padding on the FFMA path
    apply = u.apply;
    u.p_w = pi++; u.p_g = pi++; u.p_b = pi++;
    u.bn = bn++;
    u.neg_off = neg; neg += Cout;
    const size_t px = B * u.H * u.W;
    u.z = tbump(&cur, px * Cout * es);
    if (apply) u.a = tbump(&cur, px * Cout * es);
    if (pool) u.pooled = tbump(&cur, B * (u.H / 2) * (u.W / 2) * Cout * es);
    u.dz = tbump(&cur, px * Cout * es);
    if (need_dx) {
      u.din0 = tbump(&cur, px * C0 * es);
      if (C1) u.din1 = tbump(&cur, px * C1 * es);
    }
    if (f32) { u.dstat = dz_cur; dz_cur += 2 * (size_t)Cout * 8; u.dbwd = dz_cur; dz_cur += 2 * (size_t)Cout * 8; }
    u.consts = tbump(&cur, 4 * (size_t)Cout * 4);
    u.stats = zero_cur; zero_cur += align_up(2 * (size_t)Cout * 4, 256);
    const size_t ktot = 9 * (size_t)(u.C0 + C1);
    u.w_fwd = packed_cur; packed_cur += align_up(ktot * Cout * es, 256);
    if (!first) { u.w_dgrad = packed_cur; packed_cur += align_up(ktot * Cout * es, 256); }
    if (!f32) { u.dwk = dwk_cur; dwk_cur += align_up(ktot * Cout * 4, 256); }
    p->param_numel.push_back((long long)Cout * cin_real * 9);
    p->param_numel.push_back(Cout);
    p->param_numel.push_back(Cout);
    return u;
  };
  // parameters() order of the reference module (unet.py:67-77): inc, down.0.., up.0.. (up, conv), outc
  for (int l = 0; l <= p->depth; ++l) {
    const int cin = l == 0 ? g->in_channels : dims[l - 1];
    p->enc.push_back(make_unit(l, l == 0 ? 16 : cin, 0, cin, dims[l], l == 0, false, true, l > 0));
    p->enc.push_back(make_unit(l, dims[l], 0, dims[l], dims[l], false, l < p->depth, true, true));
  }
  for (int i = 0; i < p->depth; ++i) {
    const int l = p->depth - 1 - i;
    TrainUp t;
    t.Cin = dims[l + 1]; t.Cout = dims[l]; t.hs = p->Hs[l + 1]; t.ws = p->Ws[l + 1];
    t.p_w = pi++; t.p_b = pi++;
    p->param_numel.push_back(4LL * t.Cin * t.Cout);
    p->param_numel.push_back(t.Cout);
    t.u = tbump(&cur, B * (2 * t.hs) * (2 * t.ws) * t.Cout * es);
    t.din = tbump(&cur, B * t.hs * t.ws * t.Cin * es);
    t.w_fwd = packed_cur; packed_cur += align_up(4 * (size_t)t.Cin * t.Cout * es, 256);
    t.w_dgrad = packed_cur; packed_cur += align_up(4 * (size_t)t.Cin * t.Cout * es, 256);
    if (f32) { t.dbias = dz_cur; dz_cur += (size_t)t.Cout * 8; }
    t.bias4 = tbump(&cur, 4 * (size_t)t.Cout * 4);
    p->ups.push_back(t);
    TrainUnit u1 = make_unit(l, dims[l], dims[l], 2 * dims[l], dims[l], false, false, true, true);
    u1.H1 = 2 * t.hs; u1.W1 = 2 * t.ws;
    u1.off_y = (p->Hs[l] - 2 * t.hs) / 2; u1.off_x = (p->Ws[l] - 2 * t.ws) / 2;      // F.pad left/top = diff // 2 (unet.py:46-47)
    p->dec.push_back(u1);
    p->dec.push_back(make_unit(l, dims[l], 0, dims[l], dims[l], false, false, i < p->depth - 1, true));
  }
  p->param_numel.push_back((long long)g->n_classes * dims[0]);
  p->param_numel.push_back(g->n_classes);
  pi += 2;
  p->n_params = pi;
  p->n_bn = bn;
  p->neg_total = neg;
  for (int l = 0; l < p->depth; ++l) p->dfull.push_back(tbump(&cur, B * p->Hs[l] * p->Ws[l] * dims[l] * es));
  if (f32) {
    p->dhead = dz_cur; dz_cur += ((size_t)g->n_classes * dims[0] + g->n_classes) * 8;
    p->dzero_bytes = dz_cur;
    p->dzero = tbump(&cur, dz_cur);
    p->da_head = tbump(&cur, B * g->height * g->width * dims[0] * 4);
  }
  p->neg = tbump(&cur, (size_t)neg * 4);
  p->loss = zero_cur; zero_cur += 256;
  p->zero_bytes = zero_cur;
  p->zero_arena = tbump(&cur, zero_cur);
  p->packed_bytes = packed_cur;
  p->packed = tbump(&cur, packed_cur);
  p->dwk_bytes = dwk_cur;
  p->dwk0 = tbump(&cur, dwk_cur);
  p->pack_table = tbump(&cur, 64 * sizeof(gsd_pack_item));
  p->prep_fwd = tbump(&cur, 64 * sizeof(PrepItem));
  p->prep_bwd = tbump(&cur, 96 * sizeof(PrepItem));
  p->ws_bytes = align_up(cur, 1024);
  // resolve offsets that were relative to their arenas
  auto fix = [&](TrainUnit& u) {
    u.stats += p->zero_arena; u.w_fwd += p->packed; if (!u.first) u.w_dgrad += p->packed; u.dwk += p->dwk0;
    u.dstat += p->dzero; u.dbwd += p->dzero;
  };
  for (auto& u : p->enc) fix(u);
  for (auto& u : p->dec) fix(u);
  for (auto& t : p->ups) { t.w_fwd += p->packed; t.w_dgrad += p->packed; t.dbias += p->dzero; }
  p->dhead += p->dzero;
  p->loss += p->zero_arena;
  // conv inputs
  for (int l = 0; l <= p->depth; ++l) {
    p->enc[2 * l].src0 = l == 0 ? p->in16 : p->enc[2 * l - 1].pooled;
    p->enc[2 * l + 1].src0 = p->enc[2 * l].a;
  }
  for (int i = 0; i < p->depth; ++i) {
    const int l = p->depth - 1 - i;
    p->ups[i].src = i == 0 ? p->enc[2 * p->depth + 1].a : p->dec[2 * i - 1].a;
    p->dec[2 * i].src0 = p->enc[2 * l + 1].a;           // skip connection first (torch.cat([x2, x1]), unet.py:48)
    p->dec[2 * i].src1 = p->ups[i].u;
    p->dec[2 * i + 1].src0 = p->dec[2 * i].a;
  }
  // concat units: one split-output dgrad launch where the launch rules give it N >= 128 with streamed weights (small batches
  // plan N = 64 work items and the tap-streaming kernel has no split output: those keep two launches)
  for (int i = 0; i < p->depth && !f32; ++i) {
    TrainUnit& u = p->dec[2 * i];
    ConvDesc d;
    void* dummy = reinterpret_cast<void*>(static_cast<uintptr_t>(256));
    d.src0 = dummy; d.C0 = u.Cout; d.B = (int)B; d.H = u.H; d.W = u.W; d.w = dummy; d.Cout = u.C0 + u.C1; d.groups = 1;
    taps3x3(&d);
    d.out = dummy; d.out2 = dummy; d.split_c = u.C0;
    if (!getenv("GSD_SPLIT_DGRAD2") && prefer_halo(d, sms)) {
      HaloLaunch L;
      const bool was = plan_only_mode();
      plan_only_mode() = true;
      const int rc = build_halo_launch(d, sms, &L);
      plan_only_mode() = was;
      u.one_dgrad = rc == 0 && L.bn >= 128 && !L.wres;
    }
  }
  if (!dry) {
    GSD_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    p->events.resize(192);
    for (auto& e : p->events) GSD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (getenv("GSD_NO_WGRAD_OVERLAP")) p->overlap = 0;
  if (dry) {               // nothing is ever dereferenced, but the tables are indexed
    p->params.assign(p->n_params, nullptr);
    p->grads.assign(p->n_params, nullptr);
    p->bnbuf.assign(2 * p->n_bn, nullptr);
  }
  *out = p;
  return 0;
}

// The same plan without a GPU: layout and step structure only.  gsd_backward on it enqueues nothing but reports gradients
// and buckets in the real order (tests/test_ddp_cpu.py drives the data-parallel host logic with it).
extern "C" int gsd_debug_train_plan_create(gsd_train_plan** out, const gsd_geometry* g) {
  dry_create_mode() = true;
  const int rc = gsd_train_plan_create(out, g, 0);
  dry_create_mode() = false;
  return rc;
}

extern "C" void gsd_train_plan_destroy(gsd_train_plan* p) {
  if (!p) return;
  for (auto e : p->events) if (e) cudaEventDestroy(e);
  if (p->side) cudaStreamDestroy(p->side);
  delete p;
}

extern "C" size_t gsd_train_plan_workspace_bytes(const gsd_train_plan* p) { return p ? p->ws_bytes : 0; }
extern "C" int gsd_train_plan_num_params(const gsd_train_plan* p) { return p ? p->n_params : 0; }
extern "C" int gsd_train_plan_num_bn(const gsd_train_plan* p) { return p ? p->n_bn : 0; }
// kernel launches one gsd_train_step enqueues (bench.py's gpu_launches claim for the training leg)
extern "C" int gsd_train_plan_launches(const gsd_train_plan* p) {
  if (!p) return 0;
  if (p->f32) return p->launches_f32;       // counted while the last step was enqueued
  const int U = (int)(p->enc.size() + p->dec.size()), D = p->depth;
  int n = 3 + (3 * U - 1) + D + 1;            // pack, constants, prologue | conv + finalize (+ apply, not the last unit) | transposed convs | head
  n += 1;                                     // MSE
  n += 1 + 2 * U + (U - 1 + D) + 2 * U;       // zero fills | BatchNorm backward (2 passes) | dgrad (concat: 2, first layer: 0) | wgrad + unpack
  for (int i = 0; i < D; ++i)
    n += 3 + ((p->dec[2 * i].H != 2 * p->ups[i].hs || p->dec[2 * i].W != 2 * p->ups[i].ws) ? 1 : 0) - (p->dec[2 * i].one_dgrad ? 1 : 0);
  n += D;                                     // max-pool backward
  return n + 2;                               // Adam + EMA, step counter
}
extern "C" int gsd_train_plan_param_numel(const gsd_train_plan* p, long long* out, int capacity) {
  GSD_CHECK(p && out && capacity >= p->n_params, "gsd_train_plan_param_numel: need room for %d entries", p ? p->n_params : 0);
  for (int i = 0; i < p->n_params; ++i) out[i] = p->param_numel[i];
  return p->n_params;
}

extern "C" int gsd_train_plan_bind(gsd_train_plan* p, const void* const* params, void* const* grads, void* const* bn_buffers,
                                   long long* const* num_batches_tracked, void* workspace) {
  GSD_CHECK(p && params && grads && bn_buffers && workspace, "gsd_train_plan_bind: null argument");
  GSD_DEVICE(p->device);
  p->ws = static_cast<char*>(workspace);
  p->params.assign(p->n_params, nullptr);
  p->grads.assign(p->n_params, nullptr);
  for (int i = 0; i < p->n_params; ++i) {
    GSD_CHECK(params[i] && grads[i], "gsd_train_plan_bind: parameter / gradient pointer %d is null", i);
    p->params[i] = static_cast<const float*>(params[i]);
    p->grads[i] = static_cast<float*>(grads[i]);
  }
  p->bnbuf.assign(2 * p->n_bn, nullptr);
  for (int i = 0; i < 2 * p->n_bn; ++i) {
    GSD_CHECK(bn_buffers[i], "gsd_train_plan_bind: BatchNorm buffer pointer %d is null", i);
    p->bnbuf[i] = static_cast<float*>(bn_buffers[i]);
  }
  p->nbt.clear();
  if (num_batches_tracked)
    for (int i = 0; i < p->n_bn; ++i) p->nbt.push_back(num_batches_tracked[i]);
  // device tables: bf16 operand pack items (one launch per step), per-step constant / zero-fill items
  std::vector<gsd_pack_item> items;
  long long units = 0;
  auto add_item = [&](int mode, const float* w, void* out, void* out_dgrad, int O, int I, int Ipad) {
    gsd_pack_item it;
    it.w = w; it.out = out; it.out_dgrad = out_dgrad; it.mode = mode; it.O = O; it.I = I; it.Ipad = Ipad; it.start = units;
    units += pack_item_units(mode, O, I, Ipad);
    items.push_back(it);
  };
  auto add_unit = [&](const TrainUnit& u) {
    if (p->f32) return;                       // the fp32 path packs its operands with its own kernels (train_plan_f32.h)
    add_item(0, p->params[u.p_w], p->ws + u.w_fwd, u.first ? nullptr : p->ws + u.w_dgrad, u.Cout, u.cin_real, u.first ? 16 : u.cin_real);
  };
  for (int l = 0; l <= p->depth; ++l) { add_unit(p->enc[2 * l]); add_unit(p->enc[2 * l + 1]); }
  for (int i = 0; i < p->depth; ++i) {
    const TrainUp& t = p->ups[i];
    if (!p->f32) {
      add_item(2, p->params[t.p_w], p->ws + t.w_fwd, nullptr, t.Cout, t.Cin, t.Cin);
      add_item(3, p->params[t.p_w], p->ws + t.w_dgrad, nullptr, t.Cout, t.Cin, t.Cin);
    }
    add_unit(p->dec[2 * i]);
    add_unit(p->dec[2 * i + 1]);
  }
  GSD_CHECK(items.size() <= 64, "gsd_train_plan_bind: more than 64 pack items");
  p->n_pack = (int)items.size();
  p->pack_units = units;
  if (!items.empty()) GSD_CUDA(cudaMemcpy(p->ws + p->pack_table, items.data(), items.size() * sizeof(gsd_pack_item), cudaMemcpyHostToDevice));
  std::vector<PrepItem> fwd, bwd;
  float* negp = wsp<float>(p, p->neg);
  auto all_units = [&](auto fn) { for (auto& u : p->enc) fn(u); for (auto& u : p->dec) fn(u); };
  all_units([&](TrainUnit& u) {
    if (p->f32) return;                       // fp32: no centring, BatchNorm gradients are written, not accumulated
    fwd.push_back(PrepItem{p->bnbuf[2 * u.bn], negp + u.neg_off, u.Cout, u.Cout, -1.f});      // centring constant = -running_mean
    bwd.push_back(PrepItem{nullptr, p->grads[u.p_g], u.Cout, 1, 0.f});                          // BatchNorm gradients are accumulated
    bwd.push_back(PrepItem{nullptr, p->grads[u.p_b], u.Cout, 1, 0.f});
  });
  for (auto& t : p->ups) {
    fwd.push_back(PrepItem{p->params[t.p_b], wsp<float>(p, t.bias4), 4 * t.Cout, t.Cout, 1.f});   // bias for each of the 4 (dy,dx) groups
    if (!p->f32) bwd.push_back(PrepItem{nullptr, p->grads[t.p_b], t.Cout, 1, 0.f});
  }
  if (!p->f32) {
    bwd.push_back(PrepItem{nullptr, p->grads[p->n_params - 2], p->g.n_classes * p->g.dims[0], 1, 0.f});
    bwd.push_back(PrepItem{nullptr, p->grads[p->n_params - 1], p->g.n_classes, 1, 0.f});
  }
  GSD_CHECK(fwd.size() <= 64 && bwd.size() <= 96, "gsd_train_plan_bind: too many per-step constant items");
  p->n_prep_fwd = (int)fwd.size();
  p->n_prep_bwd = (int)bwd.size();
  if (!fwd.empty()) GSD_CUDA(cudaMemcpy(p->ws + p->prep_fwd, fwd.data(), fwd.size() * sizeof(PrepItem), cudaMemcpyHostToDevice));
  if (!bwd.empty()) GSD_CUDA(cudaMemcpy(p->ws + p->prep_bwd, bwd.data(), bwd.size() * sizeof(PrepItem), cudaMemcpyHostToDevice));
  if (p->dwk_bytes) GSD_CUDA(cudaMemset(p->ws + p->dwk0, 0, p->dwk_bytes));      // wgrad accumulators: zero once, the unpack kernel re-zeroes them
  GSD_CUDA(cudaDeviceSynchronize());
  p->bound = true;
  return 0;
}

// bucket b = arena elements [lo[b], hi[b]) and is complete when all its parameters' gradient kernels have been launched
extern "C" int gsd_train_plan_set_buckets(gsd_train_plan* p, int n_buckets, const int* bucket_of_param, const long long* lo, const long long* hi) {
  GSD_CHECK(p && (n_buckets == 0 || (bucket_of_param && lo && hi)), "gsd_train_plan_set_buckets: null argument");
  p->bucket_of.clear(); p->bucket_lo.clear(); p->bucket_hi.clear(); p->bucket_size.clear();
  if (n_buckets == 0) return 0;
  p->bucket_size.assign(n_buckets, 0);
  for (int i = 0; i < p->n_params; ++i) {
    GSD_CHECK(bucket_of_param[i] >= 0 && bucket_of_param[i] < n_buckets, "gsd_train_plan_set_buckets: parameter %d has bucket %d", i, bucket_of_param[i]);
    p->bucket_of.push_back(bucket_of_param[i]);
    ++p->bucket_size[bucket_of_param[i]];
  }
  for (int b = 0; b < n_buckets; ++b) { p->bucket_lo.push_back(lo[b]); p->bucket_hi.push_back(hi[b]); }
  return 0;
}

// x: fp32 NCHW (batch, in_channels, H, W); y: fp32 NCHW (batch, n_classes, H, W).  Train-mode BatchNorm: batch statistics,
// running statistics updated in place (momentum 0.1, unbiased variance), num_batches_tracked += 1.
extern "C" int gsd_train_forward(gsd_train_plan* p, const float* x, float* y, void* stream) {
  GSD_CHECK(p && x && y, "gsd_train_forward: null argument");
  GSD_CHECK(p->bound, "gsd_train_forward: call gsd_train_plan_bind first");
  GSD_DEVICE(p->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->f32) return f32_forward(p, x, y, st);
  const gsd_geometry& g = p->g;
  const int B = g.batch;
  GSD_TRY(gsd_op_pack_weights_batched(wsp<gsd_pack_item>(p, p->pack_table), p->n_pack, p->pack_units, st));   // every layer's bf16 operands
  GSD_CUDA(cudaMemsetAsync(p->ws + p->zero_arena, 0, p->zero_bytes, st));
  prep_kernel<<<p->n_prep_fwd, 256, 0, st>>>(wsp<PrepItem>(p, p->prep_fwd));
  GSD_CUDA(cudaGetLastError());
  const float ones[8] = {1, 1, 1, 1, 1, 1, 1, 1}, zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  GSD_TRY(gsd_op_prologue_bf16(x, nullptr, 1, 0, B, g.in_channels, g.height, g.width, g.height, g.width, ones, zeros, p->ws + p->in16, st));
  for (auto& u : p->enc) GSD_TRY(unit_forward(p, u, st));
  for (int i = 0; i < p->depth; ++i) {
    TrainUp& t = p->ups[i];
    GSD_TRY(gsd_op_conv_auto_bf16(p->ws + t.src, t.Cin, nullptr, 0, 0, 0, 0, 0, B, t.hs, t.ws, p->ws + t.w_fwd, t.Cout, 1, 4, nullptr,
                                  wsp<float>(p, t.bias4), 0, p->ws + t.u, nullptr, nullptr, p->device, st));
    GSD_TRY(unit_forward(p, p->dec[2 * i], st));
    GSD_TRY(unit_forward(p, p->dec[2 * i + 1], st));
  }
  // OutConv reads relu(BatchNorm(z)) of the last unit straight from z (unet.py:17,54-57 in one pass)
  TrainUnit& last = p->dec.back();
  const float* consts = wsp<float>(p, last.consts);
  GSD_TRY(gsd_op_bn_relu_head_fwd(p->ws + last.z, consts, consts + last.Cout, p->params[p->n_params - 2], p->params[p->n_params - 1], g.n_classes, B,
                                  g.height, g.width, y, st));
  return 0;
}

// dy: gradient of the loss w.r.t. the network output (fp32 NCHW).  Every parameter gradient is written to the pointers
// given to gsd_train_plan_bind.  cb (or NULL) fires once per bucket (gsd_train_plan_set_buckets), on the host, right after
// the last gradient kernel of that bucket was enqueued: (user, bucket, lo, hi, main stream, side stream or NULL).
extern "C" int gsd_backward(gsd_train_plan* p, const float* dy, void* stream, gsd_bucket_cb cb, void* user) {
  GSD_CHECK(p && dy, "gsd_backward: null argument");
  GSD_CHECK(p->bound || p->dry, "gsd_backward: call gsd_train_plan_bind first");
  DeviceGuard guard(p->dry ? -1 : p->device);
  GSD_TRY(guard.status());
  p->main = static_cast<cudaStream_t>(stream);
  p->cb = cb; p->cb_user = user;
  p->pending = p->bucket_size;
  p->forked = false;
  cudaStream_t st = p->main;
  const int B = p->g.batch, depth = p->depth;
  if (p->f32 && !p->dry) {
    const int rc = f32_backward(p, dy, st);
    p->cb = nullptr;
    return rc;
  }
  if (!p->dry) {
    prep_kernel<<<p->n_prep_bwd, 256, 0, st>>>(wsp<PrepItem>(p, p->prep_bwd));
    GSD_CUDA(cudaGetLastError());
  }
  const void* da = nullptr;
  // ---- decoder, last block first
  for (int i = depth - 1; i >= 0; --i) {
    TrainUnit& u1 = p->dec[2 * i];
    TrainUnit& u2 = p->dec[2 * i + 1];
    const bool last = i == depth - 1;
    GSD_TRY(unit_backward(p, u2, last ? nullptr : da, dy, true));
    if (last) { grad_done(p, p->n_params - 1); grad_done(p, p->n_params - 2); }
    GSD_TRY(unit_backward(p, u1, p->ws + u2.din0, nullptr, true));
    TrainUp& t = p->ups[i];
    // transposed conv: only the (2hs x 2ws) window of u1.din1 at (off_y, off_x) is its output gradient, the rest is F.pad
    if (!p->dry && (u1.H != 2 * t.hs || u1.W != 2 * t.ws)) {
      const long border = (long)B * ((long)u1.H * u1.W - 4L * t.hs * t.ws) * (t.Cout / 8);
      zero_frame_kernel<<<ew_grid(border), 256, 0, st>>>(wsp<__nv_bfloat16>(p, u1.din1), B, u1.H, u1.W, t.Cout / 8, u1.off_y, u1.off_x, 2 * t.hs,
                                                         2 * t.ws);
      GSD_CUDA(cudaGetLastError());
    }
    TP_RUN(p, gsd_op_convt_dgrad_bf16(p->ws + u1.din1, t.Cout, u1.H, u1.W, u1.off_y, u1.off_x, p->ws + t.w_dgrad, t.Cin, B, t.hs, t.ws, nullptr, nullptr,
                                    p->ws + t.din, p->device, st));
    cudaStream_t sw;
    GSD_TRY(side_begin(p, &sw));
    TP_RUN(p, gsd_op_bn_bwd_reduce(p->ws + u1.din1, nullptr, nullptr, nullptr, nullptr, nullptr, (long long)B * u1.H * u1.W, t.Cout, p->grads[t.p_b], sw));
    grad_done(p, t.p_b);
    if (!p->dry) GSD_CUDA(cudaMemsetAsync(p->grads[t.p_w], 0, (size_t)4 * t.Cin * t.Cout * 4, sw));
    TP_RUN(p, gsd_op_convt_wgrad_bf16(p->ws + t.src, t.Cin, p->ws + u1.din1, t.Cout, u1.H, u1.W, u1.off_y, u1.off_x, B, t.hs, t.ws, p->grads[t.p_w],
                                    p->device, sw));
    grad_done(p, t.p_w);
    da = p->ws + t.din;
  }
  // ---- encoder, bottom up: `da` is the gradient of enc[depth]'s output
  for (int l = depth; l >= 0; --l) {
    TrainUnit& u1 = p->enc[2 * l];
    TrainUnit& u2 = p->enc[2 * l + 1];
    if (l < depth) {
      // dpool from level l+1 + the skip-connection gradient of decoder block depth-1-l
      TP_RUN(p, gsd_op_maxpool_bwd(p->ws + u2.a, p->ws + p->enc[2 * l + 2].din0, p->ws + p->dec[2 * (depth - 1 - l)].din0, B, u2.H, u2.W, u2.Cout,
                                 p->ws + p->dfull[l], st));
      da = p->ws + p->dfull[l];
    }
    GSD_TRY(unit_backward(p, u2, da, nullptr, true));
    GSD_TRY(unit_backward(p, u1, p->ws + u2.din0, nullptr, l > 0));
  }
  GSD_TRY(side_join(p));
  p->cb = nullptr;
  return 0;
}

// torch.optim.Adam(lr, betas, eps, weight_decay) with coupled L2 (train_unet.py:306,375) + torch_ema 0.3 update
// (train_unet.py:309,376) over flat fp32 arenas of n elements; counter = 2 device int64 (steps, EMA updates), advanced here
extern "C" int gsd_adam_ema_step(float* param_arena, const float* grad_arena, float* m, float* v, float* ema, long long n,
                                 const gsd_adam* hp, long long* counter, void* stream) {
  GSD_CHECK(hp, "gsd_adam_ema_step: null hyper-parameters");
  return gsd_op_adam_ema_dev(param_arena, grad_arena, m, v, ema, n, hp->lr, hp->beta1, hp->beta2, hp->eps, hp->weight_decay, hp->ema_decay,
                             counter, hp->grad_scale, stream);
}

// The whole loop body.  loss: device float, overwritten with mean((y - target)^2).  opt (or NULL: gradients only):
// flat arenas that the bound parameter / gradient pointers alias.  cb(bucket = -1) fires after backward has been
// enqueued and before the optimizer kernel: the caller makes `main stream` wait for its all-reduces there.
extern "C" int gsd_train_step(gsd_train_plan* p, const float* x, const float* target, float* loss, const gsd_optimizer_state* opt,
                              void* stream, gsd_bucket_cb cb, void* user) {
  GSD_CHECK(p && x && target && loss, "gsd_train_step: null argument");
  GSD_CHECK(p->bound, "gsd_train_step: call gsd_train_plan_bind first");
  GSD_DEVICE(p->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* y = wsp<float>(p, p->y);
  float* dy = wsp<float>(p, p->dy);
  GSD_TRY(gsd_train_forward(p, x, y, stream));
  float* loss_acc = wsp<float>(p, p->loss);                    // zeroed by the forward's arena memset
  const long long n = (long long)p->g.batch * p->g.n_classes * p->g.height * p->g.width;
  GSD_TRY(gsd_op_mse(y, target, n, loss_acc, dy, stream));
  GSD_CUDA(cudaMemcpyAsync(loss, loss_acc, sizeof(float), cudaMemcpyDeviceToDevice, st));
  GSD_TRY(gsd_backward(p, dy, stream, cb, user));
  if (cb) cb(user, -1, 0, 0, stream, nullptr);
  if (opt) GSD_TRY(gsd_adam_ema_step(opt->params, opt->grads, opt->m, opt->v, opt->ema, opt->n, &opt->hp, opt->counter, stream));
  return 0;
}
