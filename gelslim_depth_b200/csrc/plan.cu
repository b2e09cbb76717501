// libgsd_b200.so -- plan construction, weight packing and the C ABI (include/gsd_b200.h).
#include <math.h>

#include <vector>

#include "../../include/gsd_b200.h"
#include "conv_fp32.cuh"
#include "conv_host.h"
#include "elementwise.cuh"
#include "host_util.h"
#include "pack_all.cuh"

using namespace gsd;

namespace {

struct ConvW {          // one packed conv3x3 (or the 2x2 transposed conv)
  size_t w_off = 0, scale_off = 0, shift_off = 0;
  int cin = 0, cin_pad = 0, cout = 0, taps = 0;
};

struct AnyLaunch {      // one GEMM launch of the network: tap-streaming kernel or halo-resident kernel
  int halo = 0;
  ConvLaunch tc;
  HaloLaunch hl;
  double flops = 0;
};

struct F32Step {        // fp32 parity mode: one conv GEMM or one 2x2 max-pool
  int pool = 0;
  F32Conv c;
  const float* pin = nullptr; float* pout = nullptr; int B = 0, H = 0, W = 0, C = 0;
  double flops = 0;
};

struct ChunkLaunches {
  int b0 = 0, nb = 0;
  std::vector<AnyLaunch> convs;
  std::vector<F32Step> f32;
  int has_first = 0;          // convs[0] also exists as the prologue-fused first conv (conv_first.cuh)
  FirstLaunch first;
};

}  // namespace

constexpr int kHostSlots = GSD_MAX_HOST_SLOTS;

struct gsd_plan {
  gsd_geometry g{};
  int device = 0, num_sms = 148;
  int depth = 0;
  std::vector<int> Hs, Ws;
  // packed arena
  std::vector<ConvW> enc;     // inc.0, inc.3, down.i.0, down.i.3 ...   (2*(depth+1))
  std::vector<ConvW> upT;     // up.i.up                                 (depth)
  std::vector<ConvW> dec;     // up.i.conv.0, up.i.conv.3               (2*depth)
  size_t head_w_off = 0, head_b_off = 0, packed_bytes = 0;
  // workspace (element offsets are in bytes)
  size_t in16_off = 0, head_tmp_off = 0, ws_bytes = 0;
  std::vector<size_t> a_off, s_off, p_off, u_off, da_off, db_off;
  // bound state
  int chunk = 0;
  int chunk_first = 0, chunk_last = 0;   // host pipeline ramp: smaller first / last chunk (0 = same as chunk)
  void* bound_ws = nullptr;
  const void* bound_packed = nullptr;
  std::vector<ChunkLaunches> chunks;
  // host-pipelined forward
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  cudaEvent_t ev_start = nullptr;
  // gsd_forward_host_async: per staging slot, "compute has consumed x_dev" and "y_host is complete"
  cudaEvent_t ev_slot_compute[kHostSlots] = {}, ev_slot_out[kHostSlots] = {};
  bool slot_used[kHostSlots] = {};
  double conv_flops = 0;
  int head_fused = 0;          // 1x1 head + de-normalisation folded into the last conv's epilogue
  int first_fused = 1;         // the last forward ran the input prologue inside the first conv (no prologue launch)
};

static size_t bump(size_t* cur, size_t bytes) {
  size_t off = align_up(*cur, 1024);
  *cur = off + bytes;
  return off;
}

extern "C" int gsd_abi_version(void) { return GSD_ABI_VERSION; }
extern "C" const char* gsd_last_error(void) { return last_error_ref().c_str(); }

extern "C" int gsd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

extern "C" int gsd_plan_create(gsd_plan** out, const gsd_geometry* g, int device) {
  GSD_CHECK(out && g, "gsd_plan_create: null argument");
  GSD_CHECK(g->dtype == GSD_DTYPE_BF16 || g->dtype == GSD_DTYPE_FP32, "gsd_plan_create: unknown dtype %d", g->dtype);
  GSD_CHECK(g->mode == GSD_MODE_INFER, "gsd_plan_create: mode %d not implemented (inference only in this build)", g->mode);
  GSD_CHECK(g->batch >= 1 && g->height >= 1 && g->width >= 1, "gsd_plan_create: bad shape");
  GSD_CHECK(g->in_channels >= 1 && g->in_channels <= 8, "gsd_plan_create: in_channels %d not in 1..8", g->in_channels);
  GSD_CHECK(g->n_classes >= 1 && g->n_classes <= 4, "gsd_plan_create: n_classes %d not in 1..4", g->n_classes);
  GSD_CHECK(g->n_dims >= 2 && g->n_dims <= GSD_MAX_DIMS, "gsd_plan_create: n_dims %d not in 2..%d", g->n_dims, GSD_MAX_DIMS);
  for (int i = 0; i < g->n_dims; ++i)
    GSD_CHECK(g->dims[i] > 0 && g->dims[i] % 64 == 0, "gsd_plan_create: layer_dimensions[%d]=%d must be a multiple of 64", i, g->dims[i]);
  for (int i = 0; i + 1 < g->n_dims; ++i)
    GSD_CHECK(g->dims[i + 1] == 2 * g->dims[i],
              "gsd_plan_create: layer_dimensions[%d]=%d must be 2x layer_dimensions[%d] (torch.cat in Up, unet.py:48)", i + 1,
              g->dims[i + 1], i);
  GSD_CHECK(g->dims[0] == 64, "gsd_plan_create: layer_dimensions[0] must be 64 (1x1 head kernel)");
  int ndev = 0;
  GSD_CUDA(cudaGetDeviceCount(&ndev));
  GSD_CHECK(device >= 0 && device < ndev, "gsd_plan_create: device %d out of range (%d devices)", device, ndev);
  int major = 0, sms = 0;
  GSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  GSD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  GSD_CHECK(major == 10, "gsd_plan_create: device %d is sm_%d0; this library contains sm_100a code only (no fallback)", device, major);

  gsd_plan* p = new gsd_plan();
  p->g = *g;
  p->device = device;
  p->num_sms = sms;
  p->depth = g->n_dims - 1;
  p->Hs.push_back(g->height);
  p->Ws.push_back(g->width);
  for (int l = 0; l < p->depth; ++l) {
    p->Hs.push_back(p->Hs.back() / 2);
    p->Ws.push_back(p->Ws.back() / 2);
  }
  if (p->Hs.back() < 1 || p->Ws.back() < 1) {
    delete p;
    return fail(-1, "gsd_plan_create: %dx%d is too small for %d poolings", g->height, g->width, g->n_dims - 1);
  }
  // ---- packed arena
  size_t cur = 0;
  const bool f32 = g->dtype == GSD_DTYPE_FP32;
  const size_t es = f32 ? 4 : 2;                 // activation / weight element size
  const int in_cpad = f32 ? g->in_channels : 16; // first-layer input channels as stored
  auto add_conv = [&](std::vector<ConvW>& v, int cin, int cin_pad, int cout, int taps, int ngroups) {
    ConvW c;
    if (f32) cin_pad = cin;
    c.cin = cin; c.cin_pad = cin_pad; c.cout = cout; c.taps = taps;
    c.w_off = bump(&cur, (size_t)ngroups * cout * taps * cin_pad * es);
    c.scale_off = bump(&cur, (size_t)ngroups * cout * 4);
    c.shift_off = bump(&cur, (size_t)ngroups * cout * 4);
    v.push_back(c);
  };
  add_conv(p->enc, g->in_channels, 16, g->dims[0], 9, 1);
  add_conv(p->enc, g->dims[0], g->dims[0], g->dims[0], 9, 1);
  for (int l = 0; l < p->depth; ++l) {
    add_conv(p->enc, g->dims[l], g->dims[l], g->dims[l + 1], 9, 1);
    add_conv(p->enc, g->dims[l + 1], g->dims[l + 1], g->dims[l + 1], 9, 1);
  }
  for (int i = 0; i < p->depth; ++i) {
    const int l = p->depth - 1 - i;
    add_conv(p->upT, g->dims[l + 1], g->dims[l + 1], g->dims[l], 1, 4);
    add_conv(p->dec, g->dims[l + 1], g->dims[l + 1], g->dims[l], 9, 1);
    add_conv(p->dec, g->dims[l], g->dims[l], g->dims[l], 9, 1);
  }
  p->head_w_off = bump(&cur, (size_t)g->n_classes * g->dims[0] * 4);
  p->head_b_off = bump(&cur, 16);
  p->packed_bytes = align_up(cur, 1024);
  // ---- workspace
  cur = 0;
  const size_t B = g->batch;
  p->in16_off = bump(&cur, B * g->height * g->width * in_cpad * es);
  for (int l = 0; l <= p->depth; ++l) {
    const size_t px = B * p->Hs[l] * p->Ws[l];
    p->a_off.push_back(bump(&cur, px * g->dims[l] * es));
    p->s_off.push_back(bump(&cur, px * g->dims[l] * es));
    if (l < p->depth) p->p_off.push_back(bump(&cur, B * p->Hs[l + 1] * p->Ws[l + 1] * g->dims[l] * es));
  }
  for (int i = 0; i < p->depth; ++i) {
    const int l = p->depth - 1 - i;
    p->u_off.push_back(bump(&cur, B * (2 * p->Hs[l + 1]) * (2 * p->Ws[l + 1]) * g->dims[l] * es));
    const size_t px = B * p->Hs[l] * p->Ws[l];
    p->da_off.push_back(bump(&cur, px * g->dims[l] * es));
    p->db_off.push_back(bump(&cur, px * g->dims[l] * es));
  }
  p->head_tmp_off = bump(&cur, B * g->n_classes * g->height * g->width * 4);
  p->ws_bytes = align_up(cur, 1024);
  p->chunk = g->batch;
  *out = p;
  return 0;
}

extern "C" void gsd_plan_destroy(gsd_plan* p) {
  if (!p) return;
  for (auto e : p->ev_in) cudaEventDestroy(e);
  for (auto e : p->ev_done) cudaEventDestroy(e);
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  for (int i = 0; i < kHostSlots; ++i) {
    if (p->ev_slot_compute[i]) cudaEventDestroy(p->ev_slot_compute[i]);
    if (p->ev_slot_out[i]) cudaEventDestroy(p->ev_slot_out[i]);
  }
  if (p->copy_in) cudaStreamDestroy(p->copy_in);
  if (p->copy_out) cudaStreamDestroy(p->copy_out);
  delete p;
}

extern "C" size_t gsd_plan_workspace_bytes(const gsd_plan* p) { return p ? p->ws_bytes : 0; }
extern "C" size_t gsd_plan_packed_bytes(const gsd_plan* p) { return p ? p->packed_bytes : 0; }
extern "C" int gsd_plan_num_params(const gsd_plan* p) { return p ? 6 * (p->depth + 1) + 8 * p->depth + 2 : 0; }
extern "C" int gsd_plan_num_bn_buffers(const gsd_plan* p) { return p ? 2 * (2 * (p->depth + 1) + 2 * p->depth) : 0; }
// Frames per chunk, in order.  gsd_forward_host pipelines H2D(c+1) | compute(c) | D2H(c-1): only the first chunk's
// upload and the last chunk's download are exposed, so those two chunks may be made smaller (the ramp).
static std::vector<int> chunk_schedule(const gsd_plan* p) {
  std::vector<int> v;
  int left = p->g.batch;
  const int last = (p->chunk_last > 0 && p->chunk_last < p->chunk) ? p->chunk_last : 0;
  if (p->chunk_first > 0 && p->chunk_first < p->chunk && left > p->chunk_first + last) {
    v.push_back(p->chunk_first);
    left -= p->chunk_first;
  }
  const int body = left - ((last && left > last) ? last : 0);
  for (int done = 0; done < body; done += p->chunk) v.push_back(body - done < p->chunk ? body - done : p->chunk);
  if (left - body > 0) v.push_back(left - body);
  return v;
}
extern "C" int gsd_plan_forward_launches(const gsd_plan* p) {
  if (!p) return 0;
  const int nchunks = (int)chunk_schedule(p).size();
  return nchunks * ((p->first_fused ? 0 : 1) + 2 * (p->depth + 1) + 3 * p->depth + (p->head_fused ? 0 : 1));   // + area resample when sizes differ
}
extern "C" int gsd_plan_set_chunk(gsd_plan* p, int frames_per_chunk) {
  GSD_CHECK(p && frames_per_chunk >= 1, "gsd_plan_set_chunk: bad argument");
  p->chunk = frames_per_chunk > p->g.batch ? p->g.batch : frames_per_chunk;
  p->chunk_first = p->chunk_last = 0;
  p->bound_ws = nullptr;   // force re-binding
  return 0;
}
extern "C" int gsd_plan_set_chunk_ramp(gsd_plan* p, int first_frames, int last_frames) {
  GSD_CHECK(p && first_frames >= 0 && last_frames >= 0, "gsd_plan_set_chunk_ramp: bad argument");
  p->chunk_first = first_frames;
  p->chunk_last = last_frames;
  p->bound_ws = nullptr;
  return 0;
}
// chunk schedule of gsd_forward_host for (batch, chunk, first, last) without a plan or a GPU (CPU tests)
extern "C" int gsd_debug_chunk_schedule(int batch, int chunk, int first, int last, int* out, int capacity) {
  GSD_CHECK(out && batch >= 1 && chunk >= 1 && first >= 0 && last >= 0, "gsd_debug_chunk_schedule: bad argument");
  gsd_plan tmp;
  tmp.g.batch = batch;
  tmp.chunk = chunk > batch ? batch : chunk;
  tmp.chunk_first = first;
  tmp.chunk_last = last;
  const std::vector<int> v = chunk_schedule(&tmp);
  GSD_CHECK((int)v.size() <= capacity, "gsd_debug_chunk_schedule: %d chunks do not fit %d slots", (int)v.size(), capacity);
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  return (int)v.size();
}
extern "C" double gsd_plan_conv_flops(const gsd_plan* p) { return p ? p->conv_flops : 0; }
extern "C" int gsd_plan_first_fused(const gsd_plan* p) { return p ? p->first_fused : 0; }

// One launch for every layer (pack_all.cuh); gate != null: only if the device-side fingerprint says the parameters changed.
static int pack_weights_impl(gsd_plan* p, const void* const* params, const void* const* bn, void* packed,
                             const unsigned long long* gate, cudaStream_t st) {
  char* base = static_cast<char*>(packed);
  const bool f32 = p->g.dtype == GSD_DTYPE_FP32;
  PackAllParams P;
  memset(&P, 0, sizeof P);
  P.eps = 1e-5f;
  int pi = 0, bi = 0, n = 0;
  long long cur = 0;
  auto F = [](const void* q) { return static_cast<const float*>(q); };
  auto push = [&](PackAllItem it, long long consts) -> int {
    GSD_CHECK(n < kPackMaxItems, "gsd_pack_weights: more than %d layers", kPackMaxItems);
    it.start = cur;
    cur += it.nw + consts;
    P.it[n++] = it;
    return 0;
  };
  auto conv_bn = [&](const ConvW& c) -> int {
    PackAllItem it = {};
    it.w = F(params[pi++]); it.a = F(params[pi++]); it.beta = F(params[pi++]);
    it.mean = F(bn[bi++]); it.var = F(bn[bi++]);
    it.out_w = base + c.w_off;
    it.scale = reinterpret_cast<float*>(base + c.scale_off); it.shift = reinterpret_cast<float*>(base + c.shift_off);
    it.kind = f32 ? 1 : 0;       // bf16: the BatchNorm scale gamma / sqrt(var + eps) is folded into the operand, only the shift is left
    it.O = c.cout; it.I = c.cin; it.Ipad = c.cin_pad; it.taps = c.taps;
    it.nw = (long long)c.cout * c.taps * c.cin_pad;
    return push(it, c.cout);
  };
  for (size_t i = 0; i < p->enc.size(); ++i) GSD_TRY(conv_bn(p->enc[i]));
  for (int i = 0; i < p->depth; ++i) {
    const ConvW& u = p->upT[i];
    PackAllItem it = {};
    it.w = F(params[pi++]); it.a = F(params[pi++]);
    it.out_w = base + u.w_off;
    it.scale = reinterpret_cast<float*>(base + u.scale_off); it.shift = reinterpret_cast<float*>(base + u.shift_off);
    it.kind = f32 ? 3 : 2; it.O = u.cout; it.I = u.cin; it.Ipad = u.cin; it.taps = 1;
    it.nw = 4LL * u.cout * u.cin;
    GSD_TRY(push(it, 4LL * u.cout));
    GSD_TRY(conv_bn(p->dec[2 * i]));
    GSD_TRY(conv_bn(p->dec[2 * i + 1]));
  }
  for (int h = 0; h < 2; ++h) {      // OutConv weight (n_classes x dims[0]) and bias: plain copies
    PackAllItem it = {};
    it.w = F(params[pi++]);
    it.out_w = base + (h == 0 ? p->head_w_off : p->head_b_off);
    it.kind = 4;
    it.nw = h == 0 ? (long long)p->g.n_classes * p->g.dims[0] : p->g.n_classes;
    GSD_TRY(push(it, 0));
  }
  P.n = n;
  P.total = cur;
  GSD_CHECK(pi == gsd_plan_num_params(p) && bi == gsd_plan_num_bn_buffers(p), "gsd_pack_weights: internal count mismatch");
  pack_all_kernel<<<148 * 8, 256, 0, st>>>(P, gate);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gsd_pack_weights(gsd_plan* p, const void* const* params, const void* const* bn, void* packed,
                                void* stream) {
  GSD_CHECK(p && params && bn && packed, "gsd_pack_weights: null argument");
  GSD_DEVICE(p->device);
  return pack_weights_impl(p, params, bn, packed, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int gsd_pack_weights_if_changed(gsd_plan* p, const void* const* params, const void* const* bn, void* packed,
                                           unsigned long long* state, void* stream) {
  GSD_CHECK(p && params && bn && packed && state, "gsd_pack_weights_if_changed: null argument");
  GSD_DEVICE(p->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FingerprintParams fp;
  memset(&fp, 0, sizeof fp);
  const int np = gsd_plan_num_params(p), nb = gsd_plan_num_bn_buffers(p);
  GSD_CHECK(np + nb <= kFpMaxTensors, "gsd_pack_weights_if_changed: more than %d tensors", kFpMaxTensors);
  // element counts in nn.Module.parameters() order, then (running_mean, running_var) per BatchNorm in module order
  long long cur = 0;
  int k = 0;
  auto add = [&](const void* ptr, long long words) {
    fp.ptr[k] = static_cast<const uint32_t*>(ptr);
    fp.start[k++] = cur;
    cur += words;
  };
  int pi = 0, bi = 0;
  std::vector<int> bn_channels;
  auto conv_bn = [&](const ConvW& c) {
    add(params[pi++], (long long)c.cout * c.cin * c.taps);
    add(params[pi++], c.cout);
    add(params[pi++], c.cout);
    bn_channels.push_back(c.cout);
  };
  for (size_t i = 0; i < p->enc.size(); ++i) conv_bn(p->enc[i]);
  for (int i = 0; i < p->depth; ++i) {
    add(params[pi++], 4LL * p->upT[i].cout * p->upT[i].cin);
    add(params[pi++], p->upT[i].cout);
    conv_bn(p->dec[2 * i]);
    conv_bn(p->dec[2 * i + 1]);
  }
  add(params[pi++], (long long)p->g.n_classes * p->g.dims[0]);
  add(params[pi++], p->g.n_classes);
  for (int c : bn_channels) { add(bn[bi++], c); add(bn[bi++], c); }
  GSD_CHECK(pi == np && bi == nb, "gsd_pack_weights_if_changed: internal count mismatch");
  fp.start[k] = cur;
  fp.n = k;
  params_fingerprint_kernel<<<148 * 8, 256, 0, st>>>(fp, state);
  GSD_CUDA(cudaGetLastError());
  return pack_weights_impl(p, params, bn, packed, state, st);
}

static void taps3x3(ConvDesc* d) {
  d->ntaps = 9;
  for (int t = 0; t < 9; ++t) { d->dy[t] = (int8_t)(t / 3 - 1); d->dx[t] = (int8_t)(t % 3 - 1); }
}

static inline int resample_grid(long rows) { return (int)(rows < 148 * 24 ? (rows < 1 ? 1 : rows) : 148 * 24); }

// (Re)build every tensor map / launch record for the given workspace + packed-weight addresses.
static int bind(gsd_plan* p, void* ws, const void* packed) {
  if (p->bound_ws == ws && p->bound_packed == packed && !p->chunks.empty()) return 0;
  p->chunks.clear();
  p->conv_flops = 0;
  const gsd_geometry& g = p->g;
  char* W = static_cast<char*>(ws);
  const char* P = static_cast<const char*>(packed);
  auto fptr = [&](size_t off) { return reinterpret_cast<const float*>(P + off); };
  // bf16 plans carry the BatchNorm scale inside the packed weights (gsd_pack_weights) and the transposed convs' scale is
  // 1: the epilogue applies no per-channel scale at all; the fp32 parity mode keeps the reference's order of operations
  auto bn_scale = [&](size_t off) { return g.dtype == GSD_DTYPE_FP32 ? fptr(off) : static_cast<const float*>(nullptr); };
  const std::vector<int> schedule = chunk_schedule(p);
  int b0 = 0;
  for (size_t ci = 0; ci < schedule.size(); b0 += schedule[ci], ++ci) {
    ChunkLaunches ch;
    ch.b0 = b0;
    ch.nb = schedule[ci];
    const size_t es = g.dtype == GSD_DTYPE_FP32 ? 4 : 2;
    auto act = [&](size_t off, int l_h, int l_w, int c) {   // address of frame b0 inside a (B,h,w,c) tensor
      return static_cast<void*>(W + off + (size_t)b0 * l_h * l_w * c * es);
    };
    auto use_halo = [&](const ConvDesc& d) {
      // halo-resident kernel wherever its fixed 16x8 tiling wastes < 25 % of the MMA rows, or the layer is one wave anyway
      return prefer_halo(d, p->num_sms);
    };
    auto add = [&](ConvDesc& d) -> int {
      if (g.dtype == GSD_DTYPE_FP32) {
        F32Step s;
        F32Conv& c = s.c;
        c.src0 = static_cast<const float*>(d.src0); c.C0 = d.C0;
        c.src1 = static_cast<const float*>(d.src1); c.C1 = d.C1; c.H1 = d.H1; c.W1 = d.W1; c.off_y = d.off_y; c.off_x = d.off_x;
        c.w = static_cast<const float*>(d.w); c.scale = d.scale; c.shift = d.shift;
        c.out = static_cast<float*>(d.out);
        c.B = d.B; c.H = d.H; c.W = d.W; c.Cout = d.Cout; c.groups = d.groups; c.ntaps = d.ntaps; c.relu = d.relu;
        for (int t = 0; t < d.ntaps; ++t) { c.dy[t] = d.dy[t]; c.dx[t] = d.dx[t]; }
        s.flops = 2.0 * d.B * d.H * d.W * (double)d.groups * d.Cout * d.ntaps * (d.C0 + d.C1);
        p->conv_flops += s.flops;
        ch.f32.push_back(s);
        if (d.pooled) {
          F32Step q;
          q.pool = 1; q.pin = c.out; q.pout = static_cast<float*>(d.pooled); q.B = d.B; q.H = d.H; q.W = d.W; q.C = d.Cout;
          ch.f32.push_back(q);
        }
        return 0;
      }
      AnyLaunch A;
      const int pdl = getenv("GSD_NO_PDL") ? 0 : 1;
      if (use_halo(d)) {
        A.halo = 1;
        // 64-channel layers are bound by the smem data pipe (operand reads + epilogue constants): their additive constant
        // goes through one extra UMMA per tile instead (bias_mma.cuh) and the epilogue loads no constants at all
        if (d.Cout == 64 && d.shift && !d.scale && !getenv("GSD_NO_BIAS_MMA")) { d.bias = d.shift; d.shift = nullptr; }
        GSD_TRY(build_halo_launch(d, p->num_sms, &A.hl));
        A.hl.pdl = pdl;
        A.flops = A.hl.flops;
      } else {
        GSD_TRY(build_conv_launch(d, p->num_sms, &A.tc));
        A.tc.pdl = pdl;
        A.flops = A.tc.flops;
      }
      p->conv_flops += A.flops;
      ch.convs.push_back(A);
      return 0;
    };
    // encoder
    for (int l = 0; l <= p->depth; ++l) {
      const int h = p->Hs[l], w = p->Ws[l];
      ConvDesc d0;
      taps3x3(&d0);
      d0.B = ch.nb; d0.H = h; d0.W = w;
      if (l == 0) { d0.C0 = g.dtype == GSD_DTYPE_FP32 ? g.in_channels : 16; d0.C0_real = g.in_channels; d0.src0 = act(p->in16_off, h, w, d0.C0); }
      else { d0.src0 = act(p->p_off[l - 1], h, w, g.dims[l - 1]); d0.C0 = g.dims[l - 1]; }
      const ConvW& c0 = p->enc[2 * l];
      d0.w = P + c0.w_off; d0.scale = bn_scale(c0.scale_off); d0.shift = fptr(c0.shift_off);
      d0.Cout = g.dims[l]; d0.relu = 1;
      d0.out = act(p->a_off[l], h, w, g.dims[l]);
      GSD_TRY(add(d0));
      if (l == 0 && g.dtype == GSD_DTYPE_BF16 && g.dims[0] == 64 && (g.in_channels == 3 || g.in_channels == 6)) {
        // the same layer with the input prologue fused into its producer warps (used when no resampling is needed)
        GSD_TRY(build_first_launch(d0.w, fptr(c0.shift_off), d0.out, ch.nb, h, w, g.in_channels, p->num_sms, &ch.first));
        ch.has_first = getenv("GSD_NO_BIAS_MMA") ? 0 : 1;      // the fused kernel has no epilogue-constant path
      }
      ConvDesc d1;
      taps3x3(&d1);
      d1.B = ch.nb; d1.H = h; d1.W = w;
      d1.src0 = d0.out; d1.C0 = g.dims[l];
      const ConvW& c1 = p->enc[2 * l + 1];
      d1.w = P + c1.w_off; d1.scale = bn_scale(c1.scale_off); d1.shift = fptr(c1.shift_off);
      d1.Cout = g.dims[l]; d1.relu = 1;
      d1.out = act(p->s_off[l], h, w, g.dims[l]);
      if (l < p->depth) d1.pooled = act(p->p_off[l], p->Hs[l + 1], p->Ws[l + 1], g.dims[l]);
      GSD_TRY(add(d1));
    }
    // decoder
    for (int i = 0; i < p->depth; ++i) {
      const int l = p->depth - 1 - i;
      const int hs = p->Hs[l + 1], wsz = p->Ws[l + 1], h = p->Hs[l], w = p->Ws[l];
      ConvDesc t;
      t.ntaps = 1; t.groups = 4;
      t.B = ch.nb; t.H = hs; t.W = wsz;
      t.src0 = (i == 0) ? act(p->s_off[p->depth], hs, wsz, g.dims[l + 1]) : act(p->db_off[i - 1], hs, wsz, g.dims[l + 1]);
      t.C0 = g.dims[l + 1];
      const ConvW& u = p->upT[i];
      t.w = P + u.w_off; t.scale = bn_scale(u.scale_off); t.shift = fptr(u.shift_off);     // transposed conv: scale is exactly 1
      t.Cout = g.dims[l]; t.relu = 0;
      t.out = act(p->u_off[i], 2 * hs, 2 * wsz, g.dims[l]);
      GSD_TRY(add(t));
      ConvDesc d0;
      taps3x3(&d0);
      d0.B = ch.nb; d0.H = h; d0.W = w;
      d0.src0 = act(p->s_off[l], h, w, g.dims[l]); d0.C0 = g.dims[l];
      d0.src1 = t.out; d0.C1 = g.dims[l]; d0.H1 = 2 * hs; d0.W1 = 2 * wsz;
      d0.off_y = (h - 2 * hs) / 2; d0.off_x = (w - 2 * wsz) / 2;       // F.pad left/top = diff // 2 (unet.py:46-47)
      const ConvW& c0 = p->dec[2 * i];
      d0.w = P + c0.w_off; d0.scale = bn_scale(c0.scale_off); d0.shift = fptr(c0.shift_off);
      d0.Cout = g.dims[l]; d0.relu = 1;
      d0.out = act(p->da_off[i], h, w, g.dims[l]);
      GSD_TRY(add(d0));
      ConvDesc d1;
      taps3x3(&d1);
      d1.B = ch.nb; d1.H = h; d1.W = w;
      d1.src0 = d0.out; d1.C0 = g.dims[l];
      const ConvW& c1 = p->dec[2 * i + 1];
      d1.w = P + c1.w_off; d1.scale = bn_scale(c1.scale_off); d1.shift = fptr(c1.shift_off);
      d1.Cout = g.dims[l]; d1.relu = 1;
      d1.out = act(p->db_off[i], h, w, g.dims[l]);
      if (i == p->depth - 1 && g.dtype == GSD_DTYPE_BF16 && use_halo(d1) && g.dims[0] == 64 && !getenv("GSD_NO_HEAD_FUSION")) {
        // OutConv + bias + denormalize_depth_image ride in this conv's epilogue; its bf16 output is never written
        d1.head_w = fptr(p->head_w_off); d1.head_b = fptr(p->head_b_off); d1.head_ncls = g.n_classes;
        d1.head_y = reinterpret_cast<float*>(W + p->head_tmp_off);   // patched per call in run_chunk
        d1.out = nullptr;
        p->head_fused = 1;
      } else if (i == p->depth - 1) {
        p->head_fused = 0;
      }
      GSD_TRY(add(d1));
    }
    p->chunks.push_back(std::move(ch));
  }
  p->bound_ws = ws;
  p->bound_packed = packed;
  return 0;
}

static int check_prepost(const gsd_plan* p, const gsd_prepost* pp, const float* base) {
  GSD_CHECK(pp != nullptr, "gsd_forward: gsd_prepost is required");
  GSD_CHECK(pp->raw_height < 65536 && pp->raw_width < 65536 && pp->out_height < 65536 && pp->out_width < 65536,
            "gsd_forward: image sides must be < 65536 (32-bit resampling arithmetic)");
  GSD_CHECK(pp->raw_height >= p->g.height && pp->raw_width >= p->g.width,
            "gsd_forward: raw frames (%dx%d) smaller than the network input (%dx%d) are not supported", pp->raw_height,
            pp->raw_width, p->g.height, p->g.width);
  GSD_CHECK(pp->out_height >= 1 && pp->out_width >= 1, "gsd_forward: bad output size");
  GSD_CHECK(!pp->use_diff || base != nullptr, "gsd_forward: use_diff set but base is NULL");
  GSD_CHECK(!pp->split_fingers || (p->g.batch % 2 == 0 && (int)p->chunks.size() <= 1 && p->chunk >= p->g.batch),
            "gsd_forward: split_fingers needs an even batch processed as one chunk");
  const int frames = pp->split_fingers ? p->g.batch / 2 : p->g.batch;
  GSD_CHECK(!pp->use_diff || pp->base_batch == 1 || pp->base_batch == frames, "gsd_forward: base_batch must be 1 or the number of frames");
  return 0;
}

// one chunk of frames through the whole network on `st`
static int run_chunk(gsd_plan* p, const ChunkLaunches& ch, const void* x, const float* base, const gsd_prepost* pp,
                     float* y, void* ws, const void* packed, cudaStream_t st, std::vector<cudaEvent_t>* evs = nullptr,
                     std::vector<double>* ev_flops = nullptr) {
  // profiling only: an event after every launch, with the 2*M*N*K of the launch that just went out (0 = memory-bound pass)
  auto mark = [&](double flops = 0.0) -> int {
    if (evs) {
      cudaEvent_t e;
      GSD_CUDA(cudaEventCreate(&e));
      GSD_CUDA(cudaEventRecord(e, st));
      evs->push_back(e);
      if (ev_flops) ev_flops->push_back(flops);
    }
    return 0;
  };
  GSD_TRY(mark());
  const gsd_geometry& g = p->g;
  char* W = static_cast<char*>(ws);
  const char* P = static_cast<const char*>(packed);
  PreParams pre;
  const size_t raw_frame = (size_t)g.in_channels * pp->raw_height * pp->raw_width;
  // frame pointers of this chunk (element size 1 for uint8 frames); the Left/Right split needs the whole batch
  const size_t esz = pp->input_u8 ? 1 : 4;
  pre.x = reinterpret_cast<const char*>(x) + (pp->split_fingers ? 0 : (size_t)ch.b0 * raw_frame * esz);
  pre.base = pp->use_diff ? ((pp->base_batch == 1 || pp->split_fingers) ? base : base + (size_t)ch.b0 * raw_frame) : nullptr;
  pre.base_batch = pp->base_batch;
  pre.use_diff = pp->use_diff;
  pre.split_fingers = pp->split_fingers;
  pre.input_u8 = pp->input_u8;
  pre.B = ch.nb; pre.C = g.in_channels; pre.Hr = pp->raw_height; pre.Wr = pp->raw_width; pre.H = g.height; pre.W = g.width;
  for (int c = 0; c < 8; ++c) { pre.in_scale[c] = pp->in_scale[c]; pre.in_shift[c] = pp->in_shift[c]; }
  const long npix = (long)g.height * g.width;
  const bool resample = pp->out_height != g.height || pp->out_width != g.width;
  float* head_out = resample ? reinterpret_cast<float*>(W + p->head_tmp_off) + (size_t)ch.b0 * g.n_classes * npix
                             : y + (size_t)ch.b0 * g.n_classes * npix;
  if (g.dtype == GSD_DTYPE_FP32) {
    // ---------------- fp32 parity mode (FFMA kernels, conv_fp32.cuh)
    float* in0 = reinterpret_cast<float*>(W + p->in16_off) + (size_t)ch.b0 * npix * g.in_channels;
    prologue_f32_kernel<<<ew_grid((long)ch.nb * npix * g.in_channels), 256, 0, st>>>(pre, in0);
    GSD_CUDA(cudaGetLastError());
    GSD_TRY(mark());
    for (const F32Step& s : ch.f32) {
      if (s.pool) {
        maxpool_f32_kernel<<<ew_grid((long)s.B * (s.H / 2) * (s.W / 2) * s.C), 256, 0, st>>>(s.pin, s.B, s.H, s.W, s.C, s.pout);
      } else {
        const long M = (long)s.c.B * s.c.H * s.c.W;
        dim3 grid((unsigned)((M + 63) / 64), (unsigned)((s.c.groups * s.c.Cout + 63) / 64));
        conv_f32_kernel<<<grid, 256, 0, st>>>(s.c);
        GSD_CUDA(cudaGetLastError());
        GSD_TRY(mark(s.flops));
      }
    }
    const float* last = reinterpret_cast<const float*>(W + p->db_off[p->depth - 1]) + (size_t)ch.b0 * npix * g.dims[0];
    head_f32_kernel<<<ew_grid(npix * ch.nb * g.n_classes), 256, 0, st>>>(last, g.dims[0], reinterpret_cast<const float*>(P + p->head_w_off),
                                                                        reinterpret_cast<const float*>(P + p->head_b_off), g.n_classes,
                                                                        pp->out_scale, pp->out_shift, npix, ch.nb, head_out);
    GSD_CUDA(cudaGetLastError());
    if (resample) {
      const long opix = (long)pp->out_height * pp->out_width;
      area_resample_kernel<<<resample_grid((long)ch.nb * g.n_classes * pp->out_height), 256, 0, st>>>(
          head_out, ch.nb * g.n_classes, g.height, g.width, pp->out_height, pp->out_width,
          y + (size_t)ch.b0 * g.n_classes * opix);
      GSD_CUDA(cudaGetLastError());
    }
    GSD_TRY(mark());
    return 0;
  }
  // The difference image / finger split / normalisation run inside the first conv's producer warps whenever the raw
  // frames already have the network's size (G1/G2); with area down-sampling (G3) a separate prologue pass writes the
  // 16-channel bf16 input first.  GSD_NO_FUSED_PROLOGUE forces the two-pass form (parity tests compare the two).
  const bool fuse_first = ch.has_first && pre.Hr == pre.H && pre.Wr == pre.W && !getenv("GSD_NO_FUSED_PROLOGUE");
  p->first_fused = fuse_first ? 1 : 0;
  if (fuse_first) {
    FirstLaunch t = ch.first;
    t.p.pre = pre;
    GSD_TRY(run_first_launch(t, st));
    GSD_TRY(mark(t.flops));
  } else {
    __nv_bfloat16* in16 = reinterpret_cast<__nv_bfloat16*>(W + p->in16_off + (size_t)ch.b0 * g.height * g.width * 16 * 2);
    const int pg = ew_grid((long)ch.nb * g.height * ((g.width + 255) / 256) * 256, 256, 148 * 32);
    if (pre.Hr == pre.H && pre.Wr == pre.W) prologue_kernel<true><<<pg, 256, 0, st>>>(pre, in16);
    else prologue_kernel<false><<<pg, 256, 0, st>>>(pre, in16);
    GSD_CUDA(cudaGetLastError());
    GSD_TRY(mark());
  }
  for (size_t li = fuse_first ? 1 : 0; li < ch.convs.size(); ++li) {
    const AnyLaunch& L = ch.convs[li];
    if (L.halo && L.hl.p.head_w) {
      HaloLaunch t = L.hl;
      t.p.head_y = head_out;
      t.p.head_scale = pp->out_scale;
      t.p.head_shift = pp->out_shift;
      GSD_TRY(run_halo_launch(t, st));
    } else if (L.halo) {
      GSD_TRY(run_halo_launch(L.hl, st));
    } else {
      GSD_TRY(run_conv_launch(L.tc, st));
    }
    GSD_TRY(mark(L.flops));
  }
  if (!p->head_fused) {
    const __nv_bfloat16* last = reinterpret_cast<const __nv_bfloat16*>(W + p->db_off[p->depth - 1] + (size_t)ch.b0 * npix * 64 * 2);
    head_kernel<64><<<ew_grid(npix * ch.nb), 256, 0, st>>>(last, reinterpret_cast<const float*>(P + p->head_w_off),
                                                           reinterpret_cast<const float*>(P + p->head_b_off), g.n_classes,
                                                           pp->out_scale, pp->out_shift, npix, ch.nb, head_out);
    GSD_CUDA(cudaGetLastError());
  }
  if (resample) {
    const long opix = (long)pp->out_height * pp->out_width;
    area_resample_kernel<<<resample_grid((long)ch.nb * g.n_classes * pp->out_height), 256, 0, st>>>(
        head_out, ch.nb * g.n_classes, g.height, g.width, pp->out_height, pp->out_width,
        y + (size_t)ch.b0 * g.n_classes * opix);
    GSD_CUDA(cudaGetLastError());
  }
  GSD_TRY(mark());
  return 0;
}

extern "C" int gsd_forward(gsd_plan* p, const void* x, const float* base, const gsd_prepost* pp, float* y,
                           void* workspace, const void* packed, void* stream) {
  GSD_CHECK(p && x && y && workspace && packed, "gsd_forward: null argument");
  GSD_TRY(check_prepost(p, pp, base));
  GSD_DEVICE(p->device);
  GSD_TRY(bind(p, workspace, packed));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (const ChunkLaunches& ch : p->chunks) GSD_TRY(run_chunk(p, ch, x, base, pp, y, workspace, packed, st));
  return 0;
}

// Enqueue upload | compute | download of one batch, chunk by chunk.  slot < 0: the blocking call (staging buffers
// ordered after everything already queued on `st`).  slot >= 0: the caller rotates staging sets, so the upload of
// call k+1 only waits for the compute that last read that slot's x_dev, and the compute only waits for the download
// that last read that slot's y_dev: consecutive calls overlap (H2D(k+1) | compute(k) | D2H(k-1)).
static int enqueue_host(gsd_plan* p, const void* x_host, const float* base, const gsd_prepost* pp, float* y_host,
                        void* x_dev, float* y_dev, void* workspace, const void* packed, cudaStream_t st, int slot) {
  if (!p->copy_in) {
    GSD_CUDA(cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking));
    GSD_CUDA(cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking));
    GSD_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
  }
  while (p->ev_in.size() < p->chunks.size()) {
    cudaEvent_t a, b;
    GSD_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    GSD_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    p->ev_in.push_back(a);
    p->ev_done.push_back(b);
  }
  const gsd_geometry& g = p->g;
  const size_t raw_frame = (size_t)g.in_channels * pp->raw_height * pp->raw_width;
  const size_t out_frame = (size_t)g.n_classes * pp->out_height * pp->out_width;
  if (slot >= 0 && p->slot_used[slot]) {
    GSD_CUDA(cudaStreamWaitEvent(p->copy_in, p->ev_slot_compute[slot], 0));   // x_dev[slot] has been consumed
    GSD_CUDA(cudaStreamWaitEvent(st, p->ev_slot_out[slot], 0));               // y_dev[slot] has been downloaded
  } else {
    // work already queued on the caller's stream (e.g. a gsd_forward reading x_dev) precedes the first copy
    GSD_CUDA(cudaEventRecord(p->ev_start, st));
    GSD_CUDA(cudaStreamWaitEvent(p->copy_in, p->ev_start, 0));
    GSD_CUDA(cudaStreamWaitEvent(p->copy_out, p->ev_start, 0));
  }
  for (size_t c = 0; c < p->chunks.size(); ++c) {
    const ChunkLaunches& ch = p->chunks[c];
    const size_t esz = pp->input_u8 ? 1 : 4;
    GSD_CUDA(cudaMemcpyAsync(static_cast<char*>(x_dev) + ch.b0 * raw_frame * esz,
                             static_cast<const char*>(x_host) + ch.b0 * raw_frame * esz, ch.nb * raw_frame * esz,
                             cudaMemcpyHostToDevice, p->copy_in));
    GSD_CUDA(cudaEventRecord(p->ev_in[c], p->copy_in));
    GSD_CUDA(cudaStreamWaitEvent(st, p->ev_in[c], 0));
    GSD_TRY(run_chunk(p, ch, x_dev, base, pp, y_dev, workspace, packed, st));
    GSD_CUDA(cudaEventRecord(p->ev_done[c], st));
    GSD_CUDA(cudaStreamWaitEvent(p->copy_out, p->ev_done[c], 0));
    GSD_CUDA(cudaMemcpyAsync(y_host + ch.b0 * out_frame, y_dev + ch.b0 * out_frame, ch.nb * out_frame * 4,
                             cudaMemcpyDeviceToHost, p->copy_out));
  }
  if (slot >= 0) {
    if (!p->ev_slot_compute[slot]) {
      GSD_CUDA(cudaEventCreateWithFlags(&p->ev_slot_compute[slot], cudaEventDisableTiming));
      GSD_CUDA(cudaEventCreateWithFlags(&p->ev_slot_out[slot], cudaEventDisableTiming));
    }
    GSD_CUDA(cudaEventRecord(p->ev_slot_compute[slot], st));
    GSD_CUDA(cudaEventRecord(p->ev_slot_out[slot], p->copy_out));
    p->slot_used[slot] = true;
  }
  return 0;
}

extern "C" int gsd_forward_host(gsd_plan* p, const void* x_host, const float* base, const gsd_prepost* pp,
                                float* y_host, void* x_dev, float* y_dev, void* workspace, const void* packed,
                                void* stream) {
  GSD_CHECK(p && x_host && y_host && x_dev && y_dev && workspace && packed, "gsd_forward_host: null argument");
  GSD_TRY(check_prepost(p, pp, base));
  GSD_DEVICE(p->device);
  GSD_TRY(bind(p, workspace, packed));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GSD_TRY(enqueue_host(p, x_host, base, pp, y_host, x_dev, y_dev, workspace, packed, st, -1));
  GSD_CUDA(cudaStreamSynchronize(p->copy_out));
  GSD_CUDA(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int gsd_forward_host_async(gsd_plan* p, const void* x_host, const float* base, const gsd_prepost* pp,
                                      float* y_host, void* x_dev, float* y_dev, void* workspace, const void* packed,
                                      void* stream, int slot) {
  GSD_CHECK(p && x_host && y_host && x_dev && y_dev && workspace && packed, "gsd_forward_host_async: null argument");
  GSD_CHECK(slot >= 0 && slot < kHostSlots, "gsd_forward_host_async: slot %d out of range [0, %d)", slot, kHostSlots);
  GSD_TRY(check_prepost(p, pp, base));
  GSD_DEVICE(p->device);
  GSD_TRY(bind(p, workspace, packed));
  return enqueue_host(p, x_host, base, pp, y_host, x_dev, y_dev, workspace, packed, static_cast<cudaStream_t>(stream),
                      slot);
}

extern "C" int gsd_forward_host_wait(gsd_plan* p, int slot) {
  GSD_CHECK(p, "gsd_forward_host_wait: null plan");
  GSD_CHECK(slot >= 0 && slot < kHostSlots, "gsd_forward_host_wait: slot %d out of range [0, %d)", slot, kHostSlots);
  GSD_CHECK(p->slot_used[slot], "gsd_forward_host_wait: slot %d has no call in flight", slot);
  GSD_DEVICE(p->device);
  GSD_CUDA(cudaEventSynchronize(p->ev_slot_out[slot]));
  return 0;
}

extern "C" int gsd_op_conv_bf16(const void* src0, int C0, const void* src1, int C1, int H1, int W1, int off_y,
                                int off_x, int B, int H, int W, const void* w, int Cout, int ntaps,
                                const int8_t* tap_dy, const int8_t* tap_dx, int out_groups, const float* scale,
                                const float* shift, int relu, void* out, void* pooled, int block_n, int device,
                                void* stream) {
  GSD_CHECK(src0 && w && scale && shift && out && tap_dy && tap_dx, "gsd_op_conv_bf16: null argument");
  GSD_DEVICE(device);
  int sms = 0, major = 0;
  GSD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  GSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  GSD_CHECK(major == 10, "gsd_op_conv_bf16: device %d is not sm_100 (no fallback)", device);
  ConvDesc d;
  d.src0 = src0; d.C0 = C0; d.src1 = src1; d.C1 = src1 ? C1 : 0; d.H1 = H1; d.W1 = W1; d.off_y = off_y; d.off_x = off_x;
  d.B = B; d.H = H; d.W = W; d.w = w; d.Cout = Cout; d.groups = out_groups; d.ntaps = ntaps;
  GSD_CHECK(ntaps >= 1 && ntaps <= kMaxTaps, "gsd_op_conv_bf16: ntaps out of range");
  for (int t = 0; t < ntaps; ++t) { d.dy[t] = tap_dy[t]; d.dx[t] = tap_dx[t]; }
  d.scale = scale; d.shift = shift; d.relu = relu; d.out = out; d.pooled = pooled; d.block_n = block_n;
  ConvLaunch L;
  GSD_TRY(build_conv_launch(d, sms, &L));
  return run_conv_launch(L, static_cast<cudaStream_t>(stream));
}

extern "C" int gsd_op_image_affine(const float* x, const float* base, int base_batch, int use_diff, int B, int Cc,
                                   int Hr, int Wr, int H, int W, const float* scale8, const float* shift8, float* out,
                                   int split_fingers, int device, void* stream) {
  GSD_CHECK(x && out && scale8 && shift8, "gsd_op_image_affine: null argument");
  GSD_CHECK(!split_fingers || B % 2 == 0, "gsd_op_image_affine: split_fingers needs an even output batch (2 x frames)");
  GSD_CHECK(!use_diff || base, "gsd_op_image_affine: use_diff without base");
  GSD_CHECK(B >= 1 && Cc >= 1 && H >= 1 && W >= 1 && Hr >= 1 && Wr >= 1, "gsd_op_image_affine: bad shape");
  GSD_DEVICE(device);
  PreParams p;
  p.x = x; p.base = use_diff ? base : nullptr; p.base_batch = base_batch; p.use_diff = use_diff;
  p.B = B; p.C = Cc; p.Hr = Hr; p.Wr = Wr; p.H = H; p.W = W; p.split_fingers = split_fingers ? 1 : 0; p.input_u8 = 0;
  for (int c = 0; c < 8; ++c) { p.in_scale[c] = scale8[c]; p.in_shift[c] = shift8[c]; }
  image_affine_kernel<<<ew_grid((long)B * Cc * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, out);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// Instrumented forward: CUDA events between consecutive launches of chunk 0.. (bench.py's live roofline).
// ms_host[i] = duration of launch i (prologue, convs in network order, head[+resample]) summed over chunks;
// flops_host[i] = 2*M*N*K of that launch (0 for the memory-bound ones).  Synchronises the stream.
extern "C" int gsd_forward_profiled(gsd_plan* p, const void* x, const float* base, const gsd_prepost* pp, float* y,
                                    void* workspace, const void* packed, void* stream, float* ms_host,
                                    double* flops_host, int capacity, int* n_out) {
  GSD_CHECK(p && x && y && workspace && packed && ms_host && flops_host && n_out, "gsd_forward_profiled: null argument");
  GSD_TRY(check_prepost(p, pp, base));
  GSD_DEVICE(p->device);
  GSD_TRY(bind(p, workspace, packed));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int per_chunk = 0;
  for (const ChunkLaunches& ch : p->chunks) {
    std::vector<cudaEvent_t> evs;
    std::vector<double> fl;
    GSD_TRY(run_chunk(p, ch, x, base, pp, y, workspace, packed, st, &evs, &fl));
    GSD_CUDA(cudaStreamSynchronize(st));
    const int n = (int)evs.size() - 1;          // intervals: one per launch (the last also covers head / resample passes)
    GSD_CHECK(capacity >= n, "gsd_forward_profiled: capacity %d < %d", capacity, n);
    if (per_chunk == 0) {
      per_chunk = n;
      for (int i = 0; i < n; ++i) { ms_host[i] = 0.f; flops_host[i] = 0.0; }
    }
    for (int i = 0; i < n && i < per_chunk; ++i) {
      float ms = 0.f;
      GSD_CUDA(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
      ms_host[i] += ms;
      flops_host[i] += fl[i + 1];
    }
    for (auto e : evs) cudaEventDestroy(e);
  }
  *n_out = per_chunk;
  return 0;
}

// ---------------------------------------------------------------- per-layer taps (parity tests)
// index: 0 .. 2(depth+1)-1 = the encoder's conv -> BN -> ReLU units in network order (inc.0, inc.3, down.0.0, ...);
// then per decoder block: transposed conv (up.i.up), up.i.conv.0, up.i.conv.3.
static int tap_lookup(const gsd_plan* p, int index, size_t* off, int* C, int* H, int* W) {
  const int n_enc = 2 * (p->depth + 1);
  GSD_CHECK(index >= 0 && index < n_enc + 3 * p->depth, "activation index %d out of range [0, %d)", index, n_enc + 3 * p->depth);
  if (index < n_enc) {
    const int l = index / 2;
    *off = (index & 1) ? p->s_off[l] : p->a_off[l];
    *C = p->g.dims[l]; *H = p->Hs[l]; *W = p->Ws[l];
    return 0;
  }
  const int i = (index - n_enc) / 3, k = (index - n_enc) % 3, l = p->depth - 1 - i;
  *C = p->g.dims[l];
  if (k == 0) { *off = p->u_off[i]; *H = 2 * p->Hs[l + 1]; *W = 2 * p->Ws[l + 1]; }
  else { *off = k == 1 ? p->da_off[i] : p->db_off[i]; *H = p->Hs[l]; *W = p->Ws[l]; }
  return 0;
}

extern "C" int gsd_debug_num_activations(const gsd_plan* p) { return p ? 2 * (p->depth + 1) + 3 * p->depth : 0; }

extern "C" int gsd_debug_activation_shape(const gsd_plan* p, int index, int* C, int* H, int* W) {
  GSD_CHECK(p && C && H && W, "gsd_debug_activation_shape: null argument");
  size_t off;
  return tap_lookup(p, index, &off, C, H, W);
}

// NHWC (bf16 or fp32, the plan's dtype) activation of layer `index` left in `workspace` by the last gsd_forward ->
// fp32 NCHW (batch, C, H, W).  The last decoder unit only exists when the 1x1 head is not fused into it
// (GSD_NO_HEAD_FUSION); plans that run in several chunks are not supported.
extern "C" int gsd_debug_read_activation(const gsd_plan* p, int index, const void* workspace, float* dst, void* stream) {
  GSD_CHECK(p && workspace && dst, "gsd_debug_read_activation: null argument");
  GSD_CHECK(p->chunks.size() == 1, "gsd_debug_read_activation: run one forward with a single chunk first");
  size_t off;
  int C, H, W;
  GSD_TRY(tap_lookup(p, index, &off, &C, &H, &W));
  GSD_CHECK(!(p->head_fused && index == gsd_debug_num_activations(p) - 1),
            "gsd_debug_read_activation: the last unit's output is never stored when the 1x1 head is fused into it");
  GSD_DEVICE(p->device);
  const long total = (long)p->g.batch * C * H * W;
  const char* src = static_cast<const char*>(workspace) + off;
  if (p->g.dtype == GSD_DTYPE_FP32)
    nhwc_to_nchw_kernel<float><<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float*>(src), p->g.batch, C, H, W, dst);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src),
                                                                                                   p->g.batch, C, H, W, dst);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

// conv3x3 (pad 1) through the halo kernel (conv_halo.cuh); same tensor conventions as gsd_op_conv_bf16.
extern "C" int gsd_op_conv3x3_halo_bf16(const void* src0, int C0, const void* src1, int C1, int H1, int W1, int off_y,
                                        int off_x, int B, int H, int W, const void* w, int Cout, const float* scale,
                                        const float* shift, int relu, void* out, void* pooled, int block_n,
                                        int base_off_mode, int device, void* stream) {
  GSD_CHECK(src0 && w && scale && shift && out, "gsd_op_conv3x3_halo_bf16: null argument");
  GSD_DEVICE(device);
  int sms = 0, major = 0;
  GSD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  GSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  GSD_CHECK(major == 10, "gsd_op_conv3x3_halo_bf16: device %d is not sm_100 (no fallback)", device);
  ConvDesc d;
  d.src0 = src0; d.C0 = C0; d.src1 = src1; d.C1 = src1 ? C1 : 0; d.H1 = H1; d.W1 = W1; d.off_y = off_y; d.off_x = off_x;
  d.B = B; d.H = H; d.W = W; d.w = w; d.Cout = Cout; d.groups = 1;
  taps3x3(&d);
  d.scale = scale; d.shift = shift; d.relu = relu; d.out = out; d.pooled = pooled; d.block_n = block_n;
  HaloLaunch L;
  (void)base_off_mode;
  GSD_TRY(build_halo_launch(d, sms, &L));
  return run_halo_launch(L, static_cast<cudaStream_t>(stream));
}

// dW[co][tap][ci] += sum_pixels dZ[.,co] * X[.+tap, ci]  (conv3x3 / pad 1 weight gradient; fp32, accumulated --
// the caller zeroes dw).  X may be the virtual concat of two NHWC sources like in gsd_op_conv_bf16.
extern "C" int gsd_op_wgrad3x3_bf16(const void* x0, int C0, const void* x1, int C1, int H1, int W1, int off_y, int off_x,
                                    const void* dz, int Cout, int B, int H, int W, float* dw, int device, void* stream) {
  GSD_CHECK(x0 && dz && dw, "gsd_op_wgrad3x3_bf16: null argument");
  GSD_DEVICE(device);
  int sms = 0, major = 0;
  GSD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  GSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  GSD_CHECK(major == 10, "gsd_op_wgrad3x3_bf16: device %d is not sm_100 (no fallback)", device);
  WgradLaunch L;
  GSD_TRY(build_wgrad_launch(x0, C0, x1 ? x1 : nullptr, x1 ? C1 : 0, H1, W1, off_y, off_x, dz, Cout, B, H, W, dw, sms, &L));
  return run_wgrad_launch(L, static_cast<cudaStream_t>(stream));
}

extern "C" int gsd_op_gaussian_blur(const float* x, int planes, int H, int W, int kernel_size, float sigma, float* out, void* stream) {
  GSD_CHECK(x && out && planes >= 1 && H >= 1 && W >= 1, "gsd_op_gaussian_blur: bad argument");
  GSD_CHECK(kernel_size >= 1 && kernel_size <= 31 && (kernel_size & 1), "gsd_op_gaussian_blur: kernel_size must be odd and <= 31");
  GSD_CHECK(kernel_size / 2 < H && kernel_size / 2 < W, "gsd_op_gaussian_blur: reflect padding needs kernel_size/2 < H, W");
  GSD_DEVICE_OF(x);
  BlurParams p;
  memset(&p, 0, sizeof p);
  p.k = kernel_size; p.planes = planes; p.H = H; p.W = W;
  if (sigma <= 0.f) sigma = 0.3f * ((kernel_size - 1) * 0.5f - 1.f) + 0.8f;        // torchvision default
  double sum = 0;
  for (int i = 0; i < kernel_size; ++i) {
    const double xx = -(kernel_size - 1) * 0.5 + i;
    p.w[i] = (float)exp(-0.5 * (xx / sigma) * (xx / sigma));
    sum += p.w[i];
  }
  for (int i = 0; i < kernel_size; ++i) p.w[i] = (float)(p.w[i] / sum);
  gaussian_blur_kernel<<<ew_grid((long)planes * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, x, out);
  GSD_CUDA(cudaGetLastError());
  return 0;
}

#include "train_plan.h"
