// Memory-bound kernels around the tensor-core convs: input prologue, 1x1 head, area resampling,
// weight packing.  All are one-pass, coalesced, 128-bit vectorised where the layout allows.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsd {

struct PreParams {
  const void* x;       // raw frames, NCHW, fp32 or uint8: (B, C, Hr, Wr), or (B/2, 2C, Hr, Wr) when split_fingers
  const float* base;   // (Bb, C or 2C, Hr, Wr) fp32 or null
  int base_batch;
  int use_diff;
  int B, C, Hr, Wr, H, W;
  int split_fingers;   // general_dataset.py:71: network batch b -> finger b / (B/2) (channels [f*C, f*C+C)) of frame b % (B/2)
  int input_u8;        // 0: float 0..255 NCHW; 1: uint8 NCHW; 2: uint8 NHWC (interleaved camera bytes, (B, Hr, Wr, C))
  float in_scale[8], in_shift[8];
};

// plane (b, c) of the raw frames / base image for network batch index b, honouring the Left/Right split
__device__ __forceinline__ long pre_plane(const PreParams& p, int b, int c, int batch_of_tensor) {
  if (!p.split_fingers) return ((long)(batch_of_tensor == 1 ? 0 : b) * p.C + c);
  const int frames = p.B >> 1, f = b / frames, n = b - f * frames;
  return ((long)(batch_of_tensor == 1 ? 0 : n) * 2 * p.C + f * p.C + c);
}
__device__ __forceinline__ float pre_load(const PreParams& p, long plane, long off) {
  if (p.input_u8 == 2) {             // interleaved camera frame: plane = frame * Ct + channel
    const int ct = p.split_fingers ? 2 * p.C : p.C;
    const long n = plane / ct;
    return (float)__ldg(static_cast<const unsigned char*>(p.x) + (n * p.Hr * p.Wr + off) * ct + (plane - n * ct));
  }
  const long i = plane * p.Hr * p.Wr + off;
  return p.input_u8 ? (float)__ldg(static_cast<const unsigned char*>(p.x) + i) : __ldg(static_cast<const float*>(p.x) + i);
}

// get_difference_image (image_utils.py:6-10) -> area down-sample (image_utils.py:12-15) ->
// normalize_tactile_image (normalization_utils.py:29-34) -> NHWC bf16 with channels padded to 16
// (the first conv's K-block).  One thread per output pixel; reads are coalesced along x per channel
// plane, the 32-byte pixel record is written as two 16-byte stores.
template <bool IDENT>   // IDENT: raw size == network size (G1/G2): no resampling loops, all 2*C loads of a pixel in flight
__global__ void __launch_bounds__(256) prologue_kernel(const PreParams p, __nv_bfloat16* __restrict__ out) {
  // one block = 256 consecutive pixels of one image row segment; 32-bit index arithmetic, no per-pixel division
  const int tiles_per_row = (p.W + 255) / 256;
  const int rows = p.B * p.H;
  constexpr bool identity = IDENT;
  for (int t = blockIdx.x; t < rows * tiles_per_row; t += gridDim.x) {
    const int r = t / tiles_per_row;                    // block-uniform
    const int x = (t - r * tiles_per_row) * 256 + threadIdx.x;
    if (x >= p.W) continue;
    const int b = r / p.H, y = r - b * p.H;
    int ys = y, ye = y + 1, xs = x, xe = x + 1;
    if (!identity) {   // adaptive-average-pool bin: [floor(i*in/out), ceil((i+1)*in/out)); 32-bit (sizes < 65536)
      const unsigned uH = p.H, uW = p.W, uHr = p.Hr, uWr = p.Wr;
      ys = (int)(((unsigned)y * uHr) / uH); ye = (int)((((unsigned)y + 1u) * uHr + uH - 1u) / uH);
      xs = (int)(((unsigned)x * uWr) / uW); xe = (int)((((unsigned)x + 1u) * uWr + uW - 1u) / uW);
    }
    const float inv = 1.0f / (float)((ye - ys) * (xe - xs));
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.f;
    if (IDENT) {
      float tv[8], bv[8];
      const long off = (long)y * p.Wr + x;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < p.C) {
          tv[c] = pre_load(p, pre_plane(p, b, c, 0), off);
          bv[c] = p.use_diff ? __ldg(p.base + pre_plane(p, b, c, p.base_batch) * p.Hr * p.Wr + off) : 0.f;
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < p.C) {
          float t = tv[c];
          if (p.use_diff) t = (t - bv[c] + 255.0f) * 0.5f;
          v[c] = p.in_scale[c] * (t * inv) + p.in_shift[c];
        }
      }
    } else
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < p.C) {
        const long plx = pre_plane(p, b, c, 0);
        const float* pb = p.base ? p.base + pre_plane(p, b, c, p.base_batch) * p.Hr * p.Wr : nullptr;
        float acc = 0.f;
        for (int yy = ys; yy < ye; ++yy)
          for (int xx = xs; xx < xe; ++xx) {
            float tv = pre_load(p, plx, (long)yy * p.Wr + xx);
            if (p.use_diff) tv = (tv - __ldg(pb + (long)yy * p.Wr + xx) + 255.0f) * 0.5f;
            acc += tv;
          }
        v[c] = p.in_scale[c] * (acc * inv) + p.in_shift[c];
      }
    }
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    uint4* o = reinterpret_cast<uint4*>(out + ((long)r * p.W + x) * 16);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// fp32 parity mode: same arithmetic, NHWC fp32 with exactly C channels.
__global__ void __launch_bounds__(256) prologue_f32_kernel(const PreParams p, float* __restrict__ out) {
  const long total = (long)p.B * p.H * p.W * p.C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % p.C);
    const int x = (int)((idx / p.C) % p.W);
    const int y = (int)((idx / ((long)p.C * p.W)) % p.H);
    const int b = (int)(idx / ((long)p.C * p.W * p.H));
    const int ys = (int)(((long)y * p.Hr) / p.H), ye = (int)((((long)y + 1) * p.Hr + p.H - 1) / p.H);
    const int xs = (int)(((long)x * p.Wr) / p.W), xe = (int)((((long)x + 1) * p.Wr + p.W - 1) / p.W);
    const long plx = pre_plane(p, b, c, 0);
    const float* pb = p.base ? p.base + pre_plane(p, b, c, p.base_batch) * p.Hr * p.Wr : nullptr;
    float acc = 0.f;
    for (int yy = ys; yy < ye; ++yy)
      for (int xx = xs; xx < xe; ++xx) {
        float t = pre_load(p, plx, (long)yy * p.Wr + xx);
        if (p.use_diff) t = (t - __ldg(pb + (long)yy * p.Wr + xx) + 255.0f) * 0.5f;
        acc += t;
      }
    const int cc = c < 8 ? c : 7;
    out[idx] = p.in_scale[cc] * (acc / (float)((ye - ys) * (xe - xs))) + p.in_shift[cc];
  }
}

// OutConv 1x1 + bias (unet.py:54) + denormalize_depth_image (normalization_utils.py:129):
// (B,H,W,64) bf16 NHWC -> (B,ncls,H,W) fp32 NCHW.  Eight lanes share a pixel (16 bytes each: a warp load is 512
// contiguous bytes), a warp covers 32 pixels in 8 passes with all 8 loads in flight; the 8-lane partial dot
// products are shuffle-reduced and routed so that lane L ends up with pixel L: one coalesced 128-byte store per class.
template <int CIN>
__global__ void __launch_bounds__(256) head_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                                                   const float* __restrict__ bias, int ncls, float out_scale,
                                                   float out_shift, long npix_per_img, int B, float* __restrict__ y,
                                                   const float* __restrict__ bn_scale = nullptr,
                                                   const float* __restrict__ bn_shift = nullptr) {
  // bn_scale/bn_shift (training): `in` is the raw conv output z and the head reads relu(z*scale + shift) rounded to
  // bf16 -- exactly the tensor bn_relu_apply would have stored, which is then never materialised.
  static_assert(CIN == 64, "head_kernel: the 1x1 head reads 64 channels");
  const int lane = threadIdx.x & 31, grp = lane >> 3, c8 = lane & 7;
  float wr[4][8], bk[4], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = bn_scale ? bn_scale[c8 * 8 + j] : 1.f; sh[j] = bn_scale ? bn_shift[c8 * 8 + j] : 0.f; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bk[k] = k < ncls ? bias[k] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = k < ncls ? w[k * CIN + c8 * 8 + j] : 0.f;
  }
  const long total = npix_per_img * B;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long base = (((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < total; base += nwarps * 32) {
    uint4 u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long px = base + 4 * j + grp;
      u[j] = px < total ? __ldg(reinterpret_cast<const uint4*>(in + px * CIN + c8 * 8)) : make_uint4(0, 0, 0, 0);
    }
    float outv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t uu[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uu[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
      }
      if (bn_scale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __bfloat162float(__float2bfloat16_rn(fmaxf(f[i] * sc[i] + sh[i], 0.f)));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < ncls) {          // warp-uniform
          float sacc = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) sacc = fmaf(f[i], wr[k][i], sacc);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
          // pixel base + L = base + 4*(L>>2) + (L&3): produced in pass L>>2 by lane group L&3
          const float got = __shfl_sync(0xffffffffu, sacc, (lane & 3) * 8);
          if ((lane >> 2) == j) outv[k] = got;
        }
      }
    }
    const long px = base + lane;
    if (px < total) {
      const long b = px / npix_per_img, pix = px - b * npix_per_img;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < ncls) y[(b * ncls + k) * npix_per_img + pix] = (outv[k] + bk[k]) * out_scale + out_shift;
    }
  }
}

// F.interpolate(mode='area') == adaptive average pooling, fp32 NCHW planes (complete_prediction.py:9).
// A block walks output rows (one 64-bit division per row, none per pixel); bins of at most 2 x 2 source pixels (every
// up-sampling, e.g. the 160x213 -> 320x427 depth of the shipped pipeline) are read as four predicated loads in flight.
__global__ void __launch_bounds__(256) area_resample_kernel(const float* __restrict__ in, int planes, int Hi, int Wi,
                                                            int Ho, int Wo, float* __restrict__ out) {
  const long rows = (long)planes * Ho;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const long pl = row / Ho;
    const int y = (int)(row - pl * Ho);
    const int ys = (int)(((unsigned)y * (unsigned)Hi) / (unsigned)Ho);
    const int ye = (int)((((unsigned)y + 1u) * (unsigned)Hi + (unsigned)Ho - 1u) / (unsigned)Ho);
    const float* src = in + pl * Hi * Wi;
    float* dst = out + row * Wo;
    for (int x = threadIdx.x; x < Wo; x += 256) {
      const int xs = (int)(((unsigned)x * (unsigned)Wi) / (unsigned)Wo);
      const int xe = (int)((((unsigned)x + 1u) * (unsigned)Wi + (unsigned)Wo - 1u) / (unsigned)Wo);
      const int ny = ye - ys, nx = xe - xs;
      float acc = 0.f;
      if (ny <= 2 && nx <= 2) {
        const float* p0 = src + (long)ys * Wi + xs;
        const float v00 = __ldg(p0);
        const float v01 = nx == 2 ? __ldg(p0 + 1) : 0.f;
        const float v10 = ny == 2 ? __ldg(p0 + Wi) : 0.f;
        const float v11 = (ny == 2 && nx == 2) ? __ldg(p0 + Wi + 1) : 0.f;
        acc = v00;                                   // same accumulation order as the generic loops
        if (nx == 2) acc += v01;
        if (ny == 2) { acc += v10; if (nx == 2) acc += v11; }
      } else {
        for (int yy = ys; yy < ye; ++yy)
          for (int xx = xs; xx < xe; ++xx) acc += __ldg(src + (long)yy * Wi + xx);
      }
      dst[x] = acc / (float)(ny * nx);
    }
  }
}

// Stand-alone form of the processing helpers (image_utils.py:6-15, normalization_utils.py:4-130) for
// callers that use them outside predict_depth_from_RGB: fp32 NCHW -> fp32 NCHW,
//   out[c] = scale[c] * area_resample(use_diff ? (x - base + 255)/2 : x) + shift[c].
__global__ void __launch_bounds__(256) image_affine_kernel(const PreParams p, float* __restrict__ out) {
  const long total = (long)p.B * p.C * p.H * p.W;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % p.W);
    const int y = (int)((idx / p.W) % p.H);
    const int c = (int)((idx / ((long)p.W * p.H)) % p.C);
    const int b = (int)(idx / ((long)p.W * p.H * p.C));
    const int ys = (int)(((long)y * p.Hr) / p.H), ye = (int)((((long)y + 1) * p.Hr + p.H - 1) / p.H);
    const int xs = (int)(((long)x * p.Wr) / p.W), xe = (int)((((long)x + 1) * p.Wr + p.W - 1) / p.W);
    const long plx = pre_plane(p, b, c, 0);
    const float* pb = p.base ? p.base + pre_plane(p, b, c, p.base_batch) * p.Hr * p.Wr : nullptr;
    float acc = 0.f;
    for (int yy = ys; yy < ye; ++yy)
      for (int xx = xs; xx < xe; ++xx) {
        float t = pre_load(p, plx, (long)yy * p.Wr + xx);
        if (p.use_diff) t = (t - __ldg(pb + (long)yy * p.Wr + xx) + 255.0f) * 0.5f;
        acc += t;
      }
    const int cc = c < 8 ? c : 7;
    out[idx] = p.in_scale[cc] * (acc / (float)((ye - ys) * (xe - xs))) + p.in_shift[cc];
  }
}

// torchvision gaussian_blur (image_utils.py:17-19, general_dataset.py:76-89): depthwise k x k Gaussian with reflect
// padding on fp32 NCHW planes.  The separable 1-D weights (<= 31 taps) arrive by value; one thread per output pixel.
struct BlurParams {
  float w[32];
  int k, planes, H, W;
};
__global__ void __launch_bounds__(256) gaussian_blur_kernel(const BlurParams p, const float* __restrict__ in, float* __restrict__ out) {
  const long total = (long)p.planes * p.H * p.W;
  const int r = p.k / 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % p.W);
    const int y = (int)((idx / p.W) % p.H);
    const float* src = in + (idx / ((long)p.W * p.H)) * p.H * p.W;
    float acc = 0.f;
    for (int j = 0; j < p.k; ++j) {
      int yy = y + j - r;
      yy = yy < 0 ? -yy : (yy >= p.H ? 2 * p.H - 2 - yy : yy);          // reflect (no edge repeat)
      float row = 0.f;
      for (int i = 0; i < p.k; ++i) {
        int xx = x + i - r;
        xx = xx < 0 ? -xx : (xx >= p.W ? 2 * p.W - 2 - xx : xx);
        row = fmaf(p.w[i], __ldg(src + (long)yy * p.W + xx), row);
      }
      acc = fmaf(p.w[j], row, acc);
    }
    out[idx] = acc;
  }
}

// ---------------------------------------------------------------- weight packing
// Conv2d weight (O, I, kh, kw) fp32 -> bf16 [O][kh*kw][Ipad], zero for i >= I.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int O, int I, int taps, int Ipad,
                                        __nv_bfloat16* __restrict__ out) {
  const long total = (long)O * taps * Ipad;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % Ipad);
    const int t = (int)((idx / Ipad) % taps);
    const int o = (int)(idx / ((long)Ipad * taps));
    out[idx] = __float2bfloat16_rn(i < I ? w[((long)o * I + i) * taps + t] : 0.f);
  }
}
// The same with the eval-mode BatchNorm scale folded in: bf16(w * gamma[o] / sqrt(var[o] + eps)) -- conv(x, w*s) + shift ==
// (conv(x, w) - mean) * s + beta (unet.py:11-12), so the conv epilogue only adds (plan.cu, bf16 inference plans).
__global__ void pack_conv_weight_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ var,
                                           float eps, int O, int I, int taps, int Ipad, __nv_bfloat16* __restrict__ out) {
  const long total = (long)O * taps * Ipad;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % Ipad);
    const int t = (int)((idx / Ipad) % taps);
    const int o = (int)(idx / ((long)Ipad * taps));
    const float sc = gamma[o] * rsqrtf(var[o] + eps);          // identical expression to fold_bn_kernel's scale
    out[idx] = __float2bfloat16_rn(i < I ? w[((long)o * I + i) * taps + t] * sc : 0.f);
  }
}
// ConvTranspose2d weight (I, O, 2, 2) fp32 -> bf16 [(dy*2+dx)*O + o][I]
__global__ void pack_convt_weight_kernel(const float* __restrict__ w, int I, int O, __nv_bfloat16* __restrict__ out) {
  const long total = 4L * O * I;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const int o = (int)((idx / I) % O);
    const int g = (int)(idx / ((long)I * O));
    out[idx] = __float2bfloat16_rn(w[((long)i * O + o) * 4 + g]);
  }
}
// eval-mode BatchNorm folded to y = x*scale + shift (unet.py:12; eps 1e-5)
__global__ void fold_bn_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, int C, float eps,
                               float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = gamma[c] * rsqrtf(var[c] + eps);
    scale[c] = s;
    shift[c] = beta[c] - mean[c] * s;
  }
}
// transposed-conv epilogue constants: scale 1, shift = bias[o] for each of the 4 (dy,dx) groups
__global__ void convt_bias_kernel(const float* __restrict__ bias, int O, float* __restrict__ scale,
                                  float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * O) {
    scale[i] = 1.f;
    shift[i] = bias[i % O];
  }
}
__global__ void copy_f32_kernel(const float* __restrict__ src, long n, float* __restrict__ dst) {
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x)
    dst[idx] = src[idx];
}

// debugging taps: (B,H,W,C) activation -> fp32 (B,C,H,W)
template <class T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ in, int B, int C, int H, int W, float* __restrict__ out) {
  const long total = (long)B * C * H * W;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int c = (int)((idx / ((long)W * H)) % C);
    const long b = idx / ((long)W * H * C);
    out[idx] = (float)in[((b * H + y) * W + x) * C + c];
  }
}

inline int ew_grid(long total, int threads = 256, int cap = 148 * 16) {
  long g = (total + threads - 1) / threads;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

}  // namespace gsd
