// Memory-bound kernels of the training step (train_utils/train_unet.py:346-377): train-mode BatchNorm
// (finalize / apply / backward), fused MSE loss + gradient, 1x1 head backward, max-pool backward, reductions,
// and the fused Adam(+coupled L2)+EMA update.  NHWC bf16 activations, fp32 statistics and parameter gradients.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsd {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// BatchNorm2d in .train() (unet.py:12,15): batch mean / biased variance from the (sum, sum of squares) the conv
// epilogue accumulated; running statistics updated with momentum and the UNBIASED variance (PyTorch semantics).
// `center` (or null): the conv stored z' = z - center[c] (center = the running mean BEFORE this step): bf16 then rounds
// relative to the fluctuation of z, not to |mean| + fluctuation.  All outputs refer to the stored z'.
__global__ void bn_finalize_kernel(const float* __restrict__ stats, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float momentum, float eps, int C,
                                   const float* __restrict__ neg_center, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                   long long* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;      // BatchNorm2d bookkeeping (unet.py:12,15)
  if (c >= C) return;
  const float mean = stats[c] / count;
  float var = stats[C + c] / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float s = gamma[c] * rstd;
  const float mean_c = mean + (neg_center ? neg_center[c] : 0.f);     // batch mean of the stored (centred) tensor
  scale[c] = s;
  shift[c] = beta[c] - mean_c * s;
  mean_out[c] = mean_c;
  rstd_out[c] = rstd;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (count / fmaxf(count - 1.f, 1.f));
  }
}

__global__ void negate_f32_kernel(const float* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = -in[i];
}

// a = relu(z * scale + shift) (+ 2x2 max-pooled copy).  A block walks (image, window-row) pairs; inside a window row
// thread j owns the 2x2 window j / (C/8) x the 8-channel chunk j % (C/8).  C/8 is a power of two dividing the block
// size, so the chunk (and its scale/shift registers) never changes and the index math is shifts: the kernel is a pure
// stream of 16-byte loads and stores.
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int B, int H, int W, int C,
                                                            int c8_shift, __nv_bfloat16* __restrict__ a,
                                                            __nv_bfloat16* __restrict__ pooled) {
  const int Hw = (H + 1) / 2, Ww = (W + 1) / 2, Hp = H / 2, Wp = W / 2, C8 = C / 8;
  const int c8 = threadIdx.x & (C8 - 1);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[c8 * 8 + j]; sh[j] = shift[c8 * 8 + j]; }
  const int per_row = Ww << c8_shift;                       // (window, chunk) pairs of one window row
  const long rows = (long)B * Hw;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / Hw), wy = (int)(row - (long)b * Hw);
    const int y0 = 2 * wy;
    const bool two_rows = y0 + 1 < H;
    const __nv_bfloat16* zr = z + ((size_t)b * H + y0) * W * C;
    __nv_bfloat16* ar = a + ((size_t)b * H + y0) * W * C;
    __nv_bfloat16* pr = pooled ? pooled + ((size_t)b * Hp + wy) * Wp * C : nullptr;
    const size_t rs = (size_t)W * C;
    for (int j = threadIdx.x; j < per_row; j += 256) {
      const int wx = j >> c8_shift;
      const int x0 = 2 * wx;
      const bool two_cols = x0 + 1 < W;
      const size_t o00 = (size_t)x0 * C + c8 * 8;
      uint4 u[4];
      u[0] = *reinterpret_cast<const uint4*>(zr + o00);
      if (two_cols) u[1] = *reinterpret_cast<const uint4*>(zr + o00 + C);
      if (two_rows) {
        u[2] = *reinterpret_cast<const uint4*>(zr + rs + o00);
        if (two_cols) u[3] = *reinterpret_cast<const uint4*>(zr + rs + o00 + C);
      }
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = 0.f;               // relu output >= 0
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (((q & 1) && !two_cols) || ((q >> 1) && !two_rows)) continue;
        float f[8], r[8];
        unpack8(u[q], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i] * sc[i] + sh[i], 0.f);
        const uint4 o = pack8(f);
        *reinterpret_cast<uint4*>(ar + (q >> 1) * rs + o00 + (q & 1) * C) = o;
        unpack8(o, r);                                       // pool the ROUNDED values: pooled == max_pool2d(a) exactly
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], r[i]);
      }
      if (pr && wy < Hp && wx < Wp) *reinterpret_cast<uint4*>(pr + (size_t)wx * C + c8 * 8) = pack8(mx);
    }
  }
}

// MSE_loss (train_unet.py:51-52) and its gradient: loss += sum((y-t)^2)/n ; dy = 2 (y-t) / n
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ y, const float* __restrict__ t, long n,
                                                  float* __restrict__ loss, float* __restrict__ dy) {
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const float d = y[idx] - t[idx];
    acc += d * d;
    dy[idx] = 2.f * d * inv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
    atomicAdd(loss, s * inv);
  }
}

// OutConv backward (unet.py:54): da[pix][c] = sum_k dy[k][pix] w[k][c]; dw[k][c] += sum_pix dy a; db[k] += sum dy.
// Eight threads share a pixel (16 bytes = 8 channels each: a warp reads / writes 512 contiguous bytes) and a thread
// keeps the same 8-channel chunk for its whole pixel stride, so its 8 x ncls dW partial sums and weights live in
// registers; four pixels are in flight per thread.  Block reduction through smem atomics, then one global atomic per
// (class, channel) per block.
template <int NCLS>
__global__ void __launch_bounds__(256, 2) head_bwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ dy,
                                                       const float* __restrict__ w, unsigned npix_per_img, unsigned total,
                                                       __nv_bfloat16* __restrict__ da, float* __restrict__ dw,
                                                       float* __restrict__ db) {
  const int c8 = threadIdx.x & 7;
  const unsigned p0 = (blockIdx.x * 256u + threadIdx.x) >> 3, pstep = (gridDim.x * 256u) >> 3;
  constexpr int ncls = NCLS;
  float wr[NCLS][8], acc[NCLS][8], accb[NCLS];
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { wr[k][j] = w[k * 64 + c8 * 8 + j]; acc[k][j] = 0.f; }
  }
  for (unsigned px0 = p0; px0 < total; px0 += 4 * pstep) {
    uint4 ua[4];
    float g[4][NCLS];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned px = px0 + u * pstep;
      if (px < total) {
        ua[u] = *reinterpret_cast<const uint4*>(a + (size_t)px * 64 + c8 * 8);
        const unsigned b = px / npix_per_img, pp = px - b * npix_per_img;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) g[u][k] = __ldg(dy + ((size_t)b * ncls + k) * npix_per_img + pp);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned px = px0 + u * pstep;
      if (px >= total) break;
      float f[8], o[8];
      unpack8(ua[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) {
          v = fmaf(g[u][k], wr[k][j], v);
          acc[k][j] = fmaf(g[u][k], f[j], acc[k][j]);
        }
        o[j] = v;
      }
      *reinterpret_cast<uint4*>(da + (size_t)px * 64 + c8 * 8) = pack8(o);
      if (c8 == 0) {
#pragma unroll
        for (int k = 0; k < NCLS; ++k) accb[k] += g[u][k];
      }
    }
  }
  __shared__ float red[4 * 64 + 4];
  for (int i = threadIdx.x; i < 4 * 64 + 4; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red[k * 64 + c8 * 8 + j], acc[k][j]);
    if (c8 == 0) atomicAdd(&red[256 + k], accb[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncls * 64; i += 256) atomicAdd(dw + i, red[i]);
  if ((int)threadIdx.x < ncls) atomicAdd(db + threadIdx.x, red[256 + threadIdx.x]);
}

// Last unit of the network (up.3.conv.3 -> BatchNorm -> ReLU -> OutConv): its post-ReLU tensor `a` and the gradient
// `da` are never materialised.  Both backward passes recompute a = bf16(relu(z*scale + shift)) and
// da = bf16(sum_k dy[k] w[k][c]) from z (read once per pass) and the tiny dy planes:
//   pass 1: sums[c] = sum g, sums[C+c] = sum g*zhat  (g = da * (z*scale+shift > 0)),  dw_head += dy a,  db_head += dy
//   pass 2: dz = gamma*rstd*(g - sum_g/n - zhat*sum_gz/n)
// 3 tensor passes instead of 7 (head_bwd 2 + bn_bwd_reduce 2 + bn_bwd_apply 3).  Same thread mapping as head_bwd_kernel.
template <int NCLS, bool APPLY>
__global__ void __launch_bounds__(256, 2) head_bn_bwd_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ dy,
                                                             const float* __restrict__ w, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             float count, unsigned npix_per_img, unsigned total,
                                                             float* __restrict__ sums, float* __restrict__ dw,
                                                             float* __restrict__ db, __nv_bfloat16* __restrict__ dz,
                                                             float* __restrict__ sums2 = nullptr) {
  // sums2 (or null = sums + 64): where sum g*zhat (= dgamma) lives; separate pointers let the two reductions land straight in
  // the BatchNorm bias / weight gradient tensors
  const int c8 = threadIdx.x & 7;
  const unsigned p0 = (blockIdx.x * 256u + threadIdx.x) >> 3, pstep = (gridDim.x * 256u) >> 3;
  float wr[NCLS][8], sc[8], sh[8], mu[8], rs[8];
  float acc[NCLS][8], accb[NCLS], s1[8], s2[8];      // pass 1 accumulators
  float k1[8], k2[8], k3[8];                           // pass 2 constants
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    sc[j] = scale[c]; sh[j] = shift[c]; mu[j] = mean[c]; rs[j] = rstd[c];
    s1[j] = 0.f; s2[j] = 0.f;
    if (APPLY) { const float inv = 1.f / count; k1[j] = gamma[c] * rstd[c]; k2[j] = sums[c] * inv; k3[j] = (sums2 ? sums2[c] : sums[64 + c]) * inv; }
#pragma unroll
    for (int k = 0; k < NCLS; ++k) { wr[k][j] = w[k * 64 + c]; acc[k][j] = 0.f; }
  }
#pragma unroll
  for (int k = 0; k < NCLS; ++k) accb[k] = 0.f;
  for (unsigned px0 = p0; px0 < total; px0 += 4 * pstep) {
    uint4 uz[4];
    float g[4][NCLS];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned px = px0 + u * pstep;
      if (px < total) {
        uz[u] = *reinterpret_cast<const uint4*>(z + (size_t)px * 64 + c8 * 8);
        const unsigned b = px / npix_per_img, pp = px - b * npix_per_img;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) g[u][k] = __ldg(dy + ((size_t)b * NCLS + k) * npix_per_img + pp);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned px = px0 + u * pstep;
      if (px >= total) break;
      float zv[8], o[8];
      unpack8(uz[u], zv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = zv[j] * sc[j] + sh[j];
        float dav = 0.f;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) dav = fmaf(g[u][k], wr[k][j], dav);
        dav = __bfloat162float(__float2bfloat16_rn(dav));              // the rounding head_bwd's stored da had
        const float gg = pre > 0.f ? dav : 0.f;
        const float zh = (zv[j] - mu[j]) * rs[j];
        if (APPLY) {
          o[j] = k1[j] * (gg - k2[j] - zh * k3[j]);
        } else {
          const float av = __bfloat162float(__float2bfloat16_rn(fmaxf(pre, 0.f)));
          s1[j] += gg;
          s2[j] += gg * zh;
#pragma unroll
          for (int k = 0; k < NCLS; ++k) acc[k][j] = fmaf(g[u][k], av, acc[k][j]);
        }
      }
      if (APPLY) *reinterpret_cast<uint4*>(dz + (size_t)px * 64 + c8 * 8) = pack8(o);
      else if (c8 == 0) {
#pragma unroll
        for (int k = 0; k < NCLS; ++k) accb[k] += g[u][k];
      }
    }
  }
  if (APPLY) return;
  __shared__ float red[4 * 64 + 4 + 128];
  for (int i = threadIdx.x; i < 4 * 64 + 4 + 128; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&red[260 + c8 * 8 + j], s1[j]);
    atomicAdd(&red[260 + 64 + c8 * 8 + j], s2[j]);
#pragma unroll
    for (int k = 0; k < NCLS; ++k) atomicAdd(&red[k * 64 + c8 * 8 + j], acc[k][j]);
  }
  if (c8 == 0) {
#pragma unroll
    for (int k = 0; k < NCLS; ++k) atomicAdd(&red[256 + k], accb[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NCLS * 64; i += 256) atomicAdd(dw + i, red[i]);
  if ((int)threadIdx.x < NCLS) atomicAdd(db + threadIdx.x, red[256 + threadIdx.x]);
  if (threadIdx.x < 128)
    atomicAdd((sums2 && threadIdx.x >= 64) ? sums2 + (threadIdx.x - 64) : sums + threadIdx.x, red[260 + threadIdx.x]);
}

// BatchNorm+ReLU backward, reduction pass: sums[c] = sum g, sums[C+c] = sum g*zhat, g = da*(a>0), zhat=(z-mean)*rstd.
// With `a == nullptr` there is no ReLU (plain per-channel sum of da: transposed-conv bias gradient).
// The ReLU mask is recomputed from the stored z (z*scale + shift > 0, the very expression bn_relu_apply rounded), so
// the post-ReLU tensor is not re-read: 2 bytes/element less traffic in each backward pass.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const __nv_bfloat16* __restrict__ z,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd, long npix,
                                                            int C, long da_pix_stride, float* __restrict__ sums,
                                                            float* __restrict__ sums2 = nullptr) {
  // sums2 (or null = sums + C): destination of sum g*zhat
  // thread -> fixed 8-channel chunk; pixels strided.  blockDim.x * gridDim.x must be a multiple of C/8 (host).
  const int C8 = C / 8;
  const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int c8 = (int)(tid % C8);
  const long p0 = tid / C8, pstep = ((long)gridDim.x * blockDim.x) / C8;
  float s1[8], s2[8], mu[8], rs[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s1[j] = 0.f; s2[j] = 0.f;
    mu[j] = z ? mean[c8 * 8 + j] : 0.f; rs[j] = z ? rstd[c8 * 8 + j] : 0.f;
    sc[j] = z ? scale[c8 * 8 + j] : 0.f; sh[j] = z ? shift[c8 * 8 + j] : 0.f;
  }
  for (long px0 = p0; px0 < npix; px0 += 4 * pstep) {
    uint4 ug[4], uz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {            // issue all loads of 4 pixels first (memory-level parallelism)
      const long px = px0 + u * pstep;
      if (px < npix) {
        ug[u] = *reinterpret_cast<const uint4*>(da + px * da_pix_stride + c8 * 8);
        if (z) uz[u] = *reinterpret_cast<const uint4*>(z + px * C + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (px0 + u * pstep >= npix) break;
      float g[8];
      unpack8(ug[u], g);
      if (z) {
        float zv[8];
        unpack8(uz[u], zv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = (zv[j] * sc[j] + sh[j]) > 0.f ? g[j] : 0.f;
          s1[j] += gg;
          s2[j] += gg * (zv[j] - mu[j]) * rs[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] += g[j];
      }
    }
  }
  // block reduce per chunk through smem atomics, then one global atomic per channel per block
  extern __shared__ float smem[];   // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&smem[c8 * 8 + j], s1[j]);
    if (z) atomicAdd(&smem[C + c8 * 8 + j], s2[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
    if (smem[i] != 0.f) atomicAdd((sums2 && i >= C) ? sums2 + (i - C) : sums + i, smem[i]);
}

// BatchNorm+ReLU backward, apply pass: dz = gamma*rstd*(g - sum_g/n - zhat*sum_gz/n).
// A thread owns one 8-channel chunk for its whole pixel stride (gridDim*blockDim is a multiple of C/8), so the
// per-channel constants live in registers and the loop body is 2 loads + 1 store of 16 bytes.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ da, const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sums, float count, long npix, int C,
                                                           long da_pix_stride, __nv_bfloat16* __restrict__ dz,
                                                           const float* __restrict__ sums2 = nullptr) {
  const int C8 = C / 8;
  const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int c8 = (int)(tid % C8);
  const long p0 = tid / C8, pstep = ((long)gridDim.x * blockDim.x) / C8;
  const float inv = 1.f / count;
  float sc[8], sh[8], mu[8], rs[8], k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    sc[j] = scale[c]; sh[j] = shift[c]; mu[j] = mean[c]; rs[j] = rstd[c];
    k1[j] = gamma[c] * rstd[c];
    k2[j] = sums[c] * inv;
    k3[j] = (sums2 ? sums2[c] : sums[C + c]) * inv;
  }
  for (long px0 = p0; px0 < npix; px0 += 4 * pstep) {
    uint4 ug[4], uz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long px = px0 + u * pstep;
      if (px < npix) {
        ug[u] = *reinterpret_cast<const uint4*>(da + px * da_pix_stride + c8 * 8);
        uz[u] = *reinterpret_cast<const uint4*>(z + px * C + c8 * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long px = px0 + u * pstep;
      if (px >= npix) break;
      float g[8], zv[8], o[8];
      unpack8(ug[u], g);
      unpack8(uz[u], zv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gg = (zv[j] * sc[j] + sh[j]) > 0.f ? g[j] : 0.f;
        const float zh = (zv[j] - mu[j]) * rs[j];
        o[j] = k1[j] * (gg - k2[j] - zh * k3[j]);
      }
      *reinterpret_cast<uint4*>(dz + px * C + c8 * 8) = pack8(o);
    }
  }
}

// MaxPool2d(2) backward (unet.py:26) fused with the skip-connection gradient:
//   dfull[y,x] = dskip[y,x] (optional, pixel stride `skip_stride` elements) + (argmax of its window ? dpool : 0)
// argmax = FIRST maximum in row-major window order (PyTorch's tie rule).
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ dpool,
                                                          const __nv_bfloat16* __restrict__ dskip, long skip_stride, int B, int H,
                                                          int W, int C, int c8_shift, __nv_bfloat16* __restrict__ dfull) {
  // a block walks (image, window-row) pairs; inside a row thread j owns window j >> c8_shift x chunk j & (C/8 - 1):
  // shifts instead of 64-bit divisions, and the 4 + 1 (+ 4 skip) loads of a window are all in flight
  const int Hw = (H + 1) / 2, Ww = (W + 1) / 2, Hp = H / 2, Wp = W / 2, C8 = C / 8;
  const int c8 = threadIdx.x & (C8 - 1);
  const int per_row = Ww << c8_shift;
  const long rows = (long)B * Hw;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / Hw), wy = (int)(row - (long)b * Hw);
    for (int j = threadIdx.x; j < per_row; j += 256) {
      const int wx = j >> c8_shift;
      const bool inwin = wy < Hp && wx < Wp;
      float av[4][8], dp[8], sk[4][8];
      int arg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { arg[i] = 0; dp[i] = 0.f; }
      uint4 ua[4], us[4], ud = make_uint4(0, 0, 0, 0);
      bool live[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int y = 2 * wy + (q >> 1), x = 2 * wx + (q & 1);
        live[q] = y < H && x < W;
        const size_t pix = ((size_t)b * H + y) * W + x;
        ua[q] = (inwin) ? *reinterpret_cast<const uint4*>(a + pix * C + c8 * 8) : make_uint4(0, 0, 0, 0);
        us[q] = (live[q] && dskip) ? *reinterpret_cast<const uint4*>(dskip + pix * skip_stride + c8 * 8) : make_uint4(0, 0, 0, 0);
      }
      if (inwin) ud = *reinterpret_cast<const uint4*>(dpool + (((size_t)b * Hp + wy) * Wp + wx) * C + c8 * 8);
      if (inwin) {
        unpack8(ud, dp);
#pragma unroll
        for (int q = 0; q < 4; ++q) unpack8(ua[q], av[q]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float m = av[0][i];
#pragma unroll
          for (int q = 1; q < 4; ++q)
            if (av[q][i] > m) { m = av[q][i]; arg[i] = q; }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!live[q]) continue;
        const int y = 2 * wy + (q >> 1), x = 2 * wx + (q & 1);
        const size_t pix = ((size_t)b * H + y) * W + x;
        unpack8(us[q], sk[q]);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sk[q][i] + ((inwin && arg[i] == q) ? dp[i] : 0.f);
        *reinterpret_cast<uint4*>(dfull + pix * C + c8 * 8) = pack8(o);
      }
    }
  }
}

// conv weight for dgrad: W (O,I,3,3) fp32 -> bf16 [n = ci][tap'][k = co] with W[co][ci][8 - tap'] (flipped taps)
__global__ void pack_dgrad_weight_kernel(const float* __restrict__ w, int O, int I, __nv_bfloat16* __restrict__ out) {
  const long total = (long)I * 9 * O;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % O);
    const int t = (int)((idx / O) % 9);
    const int ci = (int)(idx / ((long)O * 9));
    out[idx] = __float2bfloat16_rn(w[((long)co * I + ci) * 9 + (8 - t)]);
  }
}
// transposed-conv weight for ITS dgrad: Wt (I,O,2,2) fp32 -> bf16 [n = ci][k = (gy, gx, co)]
__global__ void pack_convt_dgrad_weight_kernel(const float* __restrict__ w, int I, int O, __nv_bfloat16* __restrict__ out) {
  const long total = (long)I * 4 * O;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % O);
    const int g = (int)((idx / O) % 4);
    const int ci = (int)(idx / ((long)O * 4));
    out[idx] = __float2bfloat16_rn(w[((long)ci * O + co) * 4 + g]);
  }
}
// wgrad arena [O][9][Ipad] fp32 -> parameter gradient (O,I,3,3) fp32
__global__ void unpack_wgrad_kernel(float* __restrict__ dwk, int O, int I, int Ipad, float* __restrict__ grad, int clear) {
  const long total = (long)O * I * 9;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % 9);
    const int i = (int)((idx / 9) % I);
    const int o = (int)(idx / (9L * I));
    float* src = dwk + ((long)o * 9 + t) * Ipad + i;
    grad[idx] = *src;
    if (clear) *src = 0.f;          // every live element is read exactly once; padded channels are never written
  }
}

// One launch packs every layer's bf16 GEMM operands (gsd_pack_item table in device memory).  Work is counted in
// units: a Conv2d item (mode 0) has (O/32) x ceil(I/32) units, each a 32 co x 32 ci x 9 tap tile that is read once
// (288 contiguous floats per output channel), staged in shared memory and written twice -- as the forward operand
// [co][tap][ci] and, when out_dgrad is set, as the dgrad operand [ci][8-tap][co] -- with 64-byte contiguous segments
// on both sides.  ConvTranspose2d items (modes 2/3, 4 % of the parameters) are packed 1024 elements per unit.
struct PackItemDev {
  const float* w;
  __nv_bfloat16* out;
  __nv_bfloat16* out_dgrad;
  int mode, O, I, Ipad;
  long long start;       // first unit of this item
};
__device__ __forceinline__ float pack_source(const PackItemDev& it, long idx) {
  const float* w = it.w;
  switch (it.mode) {
    case 0: {   // Conv2d (O,I,3,3) -> [O][9][Ipad]
      const int i = (int)(idx % it.Ipad);
      const int t = (int)((idx / it.Ipad) % 9);
      const int o = (int)(idx / ((long)it.Ipad * 9));
      return i < it.I ? w[((long)o * it.I + i) * 9 + t] : 0.f;
    }
    case 1: {   // Conv2d dgrad operand [ci][9][co], taps flipped
      const int co = (int)(idx % it.O);
      const int t = (int)((idx / it.O) % 9);
      const int ci = (int)(idx / ((long)it.O * 9));
      return w[((long)co * it.I + ci) * 9 + (8 - t)];
    }
    case 2: {   // ConvTranspose2d (I,O,2,2) -> [(g)*O + o][I]
      const int i = (int)(idx % it.I);
      const int o = (int)((idx / it.I) % it.O);
      const int g = (int)(idx / ((long)it.I * it.O));
      return w[((long)i * it.O + o) * 4 + g];
    }
    default: {  // ConvTranspose2d dgrad operand [ci][(g, co)]
      const int co = (int)(idx % it.O);
      const int g = (int)((idx / it.O) % 4);
      const int ci = (int)(idx / ((long)it.O * 4));
      return w[((long)ci * it.O + co) * 4 + g];
    }
  }
}
__host__ __device__ inline long long pack_item_units(int mode, int O, int I, int Ipad) {
  if (mode == 0) return (long long)(O / 32) * ((I + 31) / 32);
  const long long elems = mode == 1 ? 9LL * O * I : 4LL * O * I;
  return (elems + 1023) / 1024;
}
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackItemDev* __restrict__ items, int n, long long total) {
  __shared__ long long s_start[65];
  __shared__ float tile[32][289];           // [co][ci*9 + tap], pitch 289: conflict-free for lane = ci and lane = co
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_start[i] = items[i].start;
  if (threadIdx.x == 0) s_start[n] = total;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long u = blockIdx.x; u < total; u += gridDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_start[mid] <= u) lo = mid; else hi = mid - 1;
    }
    const PackItemDev it = items[lo];
    const long long lu = u - it.start;
    if (it.mode != 0) {
      const long count = it.mode == 1 ? 9L * it.O * it.I : 4L * it.O * it.I;
      for (long idx = lu * 1024 + threadIdx.x; idx < (lu + 1) * 1024 && idx < count; idx += 256)
        it.out[idx] = __float2bfloat16_rn(pack_source(it, idx));
      continue;
    }
    const int ci_tiles = (it.I + 31) / 32;
    const int co0 = (int)(lu / ci_tiles) * 32, ci0 = (int)(lu % ci_tiles) * 32;
    const int nci = min(32, it.I - ci0);                 // real input channels in this tile
    __syncthreads();                                     // previous tile fully consumed
    for (int r = warp; r < 32; r += 8) {                 // row = output channel: 9*nci contiguous floats
      const float* src = it.w + ((long)(co0 + r) * it.I + ci0) * 9;
      for (int k = lane; k < 288; k += 32) tile[r][k] = k < 9 * nci ? src[k] : 0.f;
    }
    __syncthreads();
    // forward operand [co][tap][Ipad]: lane = ci
    const int npad = min(32, it.Ipad - ci0);             // channels to write (zero padded up to Ipad)
    for (int rt = warp; rt < 32 * 9; rt += 8) {
      const int r = rt / 9, t = rt - r * 9;
      if (lane < npad) it.out[((long)(co0 + r) * 9 + t) * it.Ipad + ci0 + lane] = __float2bfloat16_rn(tile[r][lane * 9 + t]);
    }
    // dgrad operand [ci][8 - tap][co]: lane = co
    if (it.out_dgrad) {
      for (int ct = warp; ct < nci * 9; ct += 8) {
        const int c = ct / 9, t = ct - c * 9;
        it.out_dgrad[((long)(ci0 + c) * 9 + (8 - t)) * it.O + co0 + lane] = __float2bfloat16_rn(tile[lane][c * 9 + t]);
      }
    }
  }
}

// Adam with coupled L2 (torch.optim.Adam(lr, betas, eps, weight_decay), train_unet.py:306,375) fused with the
// torch_ema==0.3 shadow update (train_unet.py:309,376), over one flat fp32 arena:
//   g += wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
//   shadow -= (1-d)*(shadow - p)
// `counter` (device, 2 x int64: Adam step t and EMA update count, BEFORE this step) makes the launch CUDA-graph
// replayable: bias corrections and the EMA warm-up decay are derived on the device; adam_tick_kernel advances it.
__global__ void adam_tick_kernel(long long* counter) {
  counter[0] += 1;
  counter[1] += 1;
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ shadow, long n, float lr,
                                                       float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                       float ema_one_minus_d, float grad_scale, const long long* __restrict__ counter,
                                                       float ema_decay, float one_minus_b1, float one_minus_b2) {
  // one_minus_b1/2 = float(1 - beta) evaluated in double on the host, as torch.optim.Adam does (`value=1 - beta2` is a Python
  // float): 1.f - 0.999f would be 1.3e-5 off
  if (counter) {
    const double t = (double)(counter[0] + 1), nu = (double)(counter[1] + 1);
    bc1 = (float)(1.0 - pow((double)b1, t));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
    double d = (1.0 + nu) / (10.0 + nu);
    if (d > (double)ema_decay) d = (double)ema_decay;
    ema_one_minus_d = (float)(1.0 - d);
  }
  const long n4 = n / 4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float4 ss = shadow ? reinterpret_cast<float4*>(shadow)[i] : make_float4(0, 0, 0, 0);
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x; float* S = &ss.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * grad_scale + wd * P[j];
      M[j] = b1 * M[j] + one_minus_b1 * gr;
      V[j] = b2 * V[j] + one_minus_b2 * gr * gr;
      P[j] -= (lr / bc1) * M[j] / (sqrtf(V[j]) / bc2_sqrt + eps);
      S[j] -= ema_one_minus_d * (S[j] - P[j]);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) reinterpret_cast<float4*>(shadow)[i] = ss;
  }
  for (long i = n4 * 4 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gr = g[i] * grad_scale + wd * p[i];
    m[i] = b1 * m[i] + one_minus_b1 * gr;
    v[i] = b2 * v[i] + one_minus_b2 * gr * gr;
    p[i] -= (lr / bc1) * m[i] / (sqrtf(v[i]) / bc2_sqrt + eps);
    if (shadow) shadow[i] -= ema_one_minus_d * (shadow[i] - p[i]);
  }
}

}  // namespace gsd
