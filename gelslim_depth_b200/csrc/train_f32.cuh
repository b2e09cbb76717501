// fp32 PARITY training path (geometry.dtype = GSD_DTYPE_FP32 on a training plan): the same step as the bf16 tensor-core
// path -- train_utils/train_unet.py:346-377 -- on fp32 NHWC activations with plain FFMA kernels, so that gradients, Adam
// moments and loss curves can be compared with the fp32 reference at fp32 tolerances (the bf16 path cannot: bf16 operand
// rounding perturbs a BatchNorm network chaotically).  About 1/50 of the bf16 path's speed; never the measured path.
// Per-channel reductions (BatchNorm statistics and their gradients, bias gradients) accumulate in double; weight gradients
// are split over pixel chunks and joined with fp32 atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_fp32.cuh"

namespace gsd {

// ---------------------------------------------------------------------------------------------- layout / packing
__global__ void __launch_bounds__(256) nchw_to_nhwc_f32_kernel(const float* __restrict__ in, int B, int C, long npix, float* __restrict__ out) {
  const long total = (long)B * C * npix;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long pix = (idx / C) % npix;
    const long b = idx / (C * npix);
    out[idx] = in[(b * C + c) * npix + pix];
  }
}

// Conv2d weight (O, I, 3, 3) -> input-gradient operand [tap'][O][I] with tap' = 8 - tap (the transposed, flipped filter)
__global__ void pack_conv_dgrad_f32_kernel(const float* __restrict__ w, int O, int I, float* __restrict__ out) {
  const long total = 9L * O * I;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const int o = (int)((idx / I) % O);
    const int t = (int)(idx / ((long)O * I));
    out[idx] = w[((long)o * I + i) * 9 + (8 - t)];
  }
}
// ConvTranspose2d weight (I, O, 2, 2) -> input-gradient operand [g][O][I]
__global__ void pack_convt_dgrad_f32_kernel(const float* __restrict__ w, int I, int O, float* __restrict__ out) {
  const long total = 4L * O * I;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % I);
    const int o = (int)((idx / I) % O);
    const int g = (int)(idx / ((long)O * I));
    out[idx] = w[((long)i * O + o) * 4 + g];
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// ---------------------------------------------------------------------------------------------- BatchNorm, train mode
// Per-channel sum / sum of squares (or, with `win`, the plain sum over a window of every image) in double.
// t: (B, H, W, C) fp32; the window is rows [oy, oy+h2) x columns [ox, ox+w2).  out[0..C) += sum, out[C..2C) += sum of squares.
__global__ void __launch_bounds__(256) chan_stats_f32_kernel(const float* __restrict__ t, int B, int H, int W, int C, int oy, int ox,
                                                             int h2, int w2, int want_sq, double* __restrict__ out) {
  const int cpb = C < 256 ? C : 256;           // channels walked by consecutive threads (coalesced NHWC rows)
  const int rows = 256 / cpb;
  const int r = threadIdx.x / cpb, c0 = threadIdx.x % cpb;
  if (r >= rows) return;
  const long npix = (long)B * h2 * w2;
  const long per = (npix + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * per, hi = lo + per < npix ? lo + per : npix;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (long pp = lo + r; pp < hi; pp += rows) {
    const int x = (int)(pp % w2), y = (int)((pp / w2) % h2);
    const long b = pp / ((long)w2 * h2);
    const float* row = t + ((b * H + oy + y) * W + ox + x) * C;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + k * cpb;
      if (c < C) {
        const double v = row[c];
        s[k] += v;
        q[k] += v * v;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k * cpb;
    if (c < C) {
      atomicAdd(out + c, s[k]);
      if (want_sq) atomicAdd(out + C + c, q[k]);
    }
  }
}

// consts = [scale | beta | mean | rstd] (scale = gamma * rstd); running statistics updated like nn.BatchNorm2d in .train()
__global__ void bn_finalize_f32_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                                       float momentum, float eps, int C, float* __restrict__ consts, long long* __restrict__ nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  consts[c] = gamma[c] * rstd;
  consts[C + c] = beta[c];
  consts[2 * C + c] = (float)mean;
  consts[3 * C + c] = rstd;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (count / (count > 1 ? count - 1 : 1)));
  }
}

// a = relu((z - mean) * (gamma * rstd) + beta)
__global__ void __launch_bounds__(256) bn_relu_apply_f32_kernel(const float* __restrict__ z, const float* __restrict__ consts, long n, int C,
                                                                float* __restrict__ a) {
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    a[idx] = fmaxf(fmaf(z[idx] - consts[2 * C + c], consts[c], consts[C + c]), 0.f);
  }
}

// sums[0..C) += sum g, sums[C..2C) += sum g * xhat with g = da where a > 0 (ReLU), xhat = (z - mean) * rstd
__global__ void __launch_bounds__(256) bn_bwd_reduce_f32_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                                                const float* __restrict__ z, const float* __restrict__ consts, long npix,
                                                                int C, double* __restrict__ sums) {
  const int cpb = C < 256 ? C : 256;
  const int rows = 256 / cpb;
  const int r = threadIdx.x / cpb, c0 = threadIdx.x % cpb;
  if (r >= rows) return;
  const long per = (npix + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * per, hi = lo + per < npix ? lo + per : npix;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (long pp = lo + r; pp < hi; pp += rows) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + k * cpb;
      if (c < C) {
        const long i = pp * C + c;
        const float g = a[i] > 0.f ? da[i] : 0.f;
        const float xh = (z[i] - consts[2 * C + c]) * consts[3 * C + c];
        s[k] += g;
        q[k] += (double)g * xh;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + k * cpb;
    if (c < C) { atomicAdd(sums + c, s[k]); atomicAdd(sums + C + c, q[k]); }
  }
}

// dz = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)); block 0 also publishes dbeta / dgamma
__global__ void __launch_bounds__(256) bn_bwd_apply_f32_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                                               const float* __restrict__ z, const float* __restrict__ consts,
                                                               const double* __restrict__ sums, double count, long n, int C,
                                                               float* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { dbeta[c] = (float)sums[c]; dgamma[c] = (float)sums[C + c]; }
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const float g = a[idx] > 0.f ? da[idx] : 0.f;
    const float xh = (z[idx] - consts[2 * C + c]) * consts[3 * C + c];
    const float mg = (float)(sums[c] / count), mgx = (float)(sums[C + c] / count);
    dz[idx] = consts[c] * (g - mg - xh * mgx);
  }
}

// ---------------------------------------------------------------------------------------------- max-pool backward
// dfull = dskip + (dpool routed to the FIRST maximum of each 2x2 window in row-major scan order, as F.max_pool2d does)
__global__ void __launch_bounds__(256) maxpool_bwd_f32_kernel(const float* __restrict__ a, const float* __restrict__ dpool,
                                                              const float* __restrict__ dskip, int B, int H, int W, int C,
                                                              float* __restrict__ dfull) {
  const int Hp = H / 2, Wp = W / 2;
  const long total = (long)B * H * W * C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int x = (int)((idx / C) % W);
    const int y = (int)((idx / ((long)C * W)) % H);
    const long b = idx / ((long)C * W * H);
    float g = dskip ? dskip[idx] : 0.f;
    const int py = y >> 1, px = x >> 1;
    if (py < Hp && px < Wp) {
      const float* s = a + ((b * H + 2 * py) * W + 2 * px) * C + c;
      const float v[4] = {s[0], s[C], s[(long)W * C], s[(long)W * C + C]};
      int arg = 0;
      float m = v[0];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k] > m) { m = v[k]; arg = k; }
      if (arg == ((y & 1) * 2 + (x & 1))) g += dpool[((b * Hp + py) * Wp + px) * C + c];
    }
    dfull[idx] = g;
  }
}

// ---------------------------------------------------------------------------------------------- OutConv backward
// a: (npix, Cin) fp32; dy: NCHW (B, ncls, npix_img).  da[p][c] = sum_k dy[k][p] w[k][c]; dw[k][c] += dy[k][p] a[p][c]; db[k] += dy[k][p]
// acc (double): [ncls * Cin | ncls]
__global__ void __launch_bounds__(256) head_bwd_f32_kernel(const float* __restrict__ a, const float* __restrict__ dy, const float* __restrict__ w,
                                                           int ncls, int Cin, long npix_img, int B, float* __restrict__ da,
                                                           double* __restrict__ acc) {
  const int cpb = Cin < 256 ? Cin : 256;
  const int rows = 256 / cpb;
  const int r = threadIdx.x / cpb, c = threadIdx.x % cpb;
  if (r >= rows) return;
  const long npix = npix_img * B;
  const long per = (npix + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * per, hi = lo + per < npix ? lo + per : npix;
  double dw[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
  float wk[4];
  for (int k = 0; k < 4; ++k) wk[k] = k < ncls ? w[k * Cin + c] : 0.f;
  for (long pp = lo + r; pp < hi; pp += rows) {
    const long b = pp / npix_img, pix = pp - b * npix_img;
    const float av = a[pp * Cin + c];
    float g = 0.f;
    for (int k = 0; k < ncls; ++k) {
      const float d = dy[(b * ncls + k) * npix_img + pix];
      g = fmaf(d, wk[k], g);
      dw[k] += (double)d * av;
      if (c == 0) db[k] += d;
    }
    da[pp * Cin + c] = g;
  }
  for (int k = 0; k < ncls; ++k) {
    atomicAdd(acc + k * Cin + c, dw[k]);
    if (c == 0) atomicAdd(acc + ncls * Cin + k, db[k]);
  }
}

// ---------------------------------------------------------------------------------------------- weight gradient
// grad[o * so + i * si + tap * st] += sum over iteration pixels (b, y, x) of  A(b, y + a_dy[tap], x + a_dx[tap])[i] *
// D(b, y * d_stride + d_dy[tap] + d_oy, x * d_stride + d_dx[tap] + d_ox)[o]     (out-of-range samples are zero)
//   Conv2d 3x3:          iteration space = output pixels, A = the conv input (virtual pad + concat of two sources), D = dz
//   ConvTranspose2d 2x2: iteration space = input pixels, A = the input, D = the output gradient at stride 2
struct F32Wgrad {
  const float* a0; int C0, aH, aW;
  const float* a1; int C1, H1, W1, off_y, off_x;
  const float* d; int Cout, dH, dW;
  int B, H, W, ntaps, d_stride, d_oy, d_ox;
  int8_t a_dy[9], a_dx[9], d_dy[9], d_dx[9];
  float* grad; long so, si, st;
  int chunk;                  // iteration pixels per block (multiple of 16)
};

__global__ void __launch_bounds__(256) wgrad_f32_kernel(const F32Wgrad p) {
  __shared__ float As[16][65];   // [pixel][input channel]
  __shared__ float Ds[16][64];   // [pixel][output channel]
  const int Ctot = p.C0 + p.C1;
  const int itiles = (Ctot + 63) / 64;
  const int tap = blockIdx.x / itiles;
  const int i0 = (blockIdx.x % itiles) * 64, o0 = blockIdx.y * 64;
  const long M = (long)p.B * p.H * p.W;
  const long m_lo = (long)blockIdx.z * p.chunk;
  const long m_hi = m_lo + p.chunk < M ? m_lo + p.chunk : M;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx -> output channels, ty -> input channels
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long m0 = m_lo; m0 < m_hi; m0 += 16) {
    const int cc = threadIdx.x & 63;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kk = (threadIdx.x >> 6) + 4 * r;
      const long m = m0 + kk;
      float av = 0.f, dv = 0.f;
      if (m < m_hi) {
        const int x = (int)(m % p.W);
        const int y = (int)((m / p.W) % p.H);
        const long b = m / ((long)p.W * p.H);
        const int i = i0 + cc;
        const int ya = y + p.a_dy[tap], xa = x + p.a_dx[tap];
        if (i < p.C0) {
          if (ya >= 0 && ya < p.aH && xa >= 0 && xa < p.aW) av = __ldg(p.a0 + ((b * p.aH + ya) * p.aW + xa) * p.C0 + i);
        } else if (i < Ctot) {
          const int y1 = ya - p.off_y, x1 = xa - p.off_x;
          if (y1 >= 0 && y1 < p.H1 && x1 >= 0 && x1 < p.W1) av = __ldg(p.a1 + ((b * p.H1 + y1) * p.W1 + x1) * p.C1 + (i - p.C0));
        }
        const int o = o0 + cc;
        const int yd = y * p.d_stride + p.d_dy[tap] + p.d_oy, xd = x * p.d_stride + p.d_dx[tap] + p.d_ox;
        if (o < p.Cout && yd >= 0 && yd < p.dH && xd >= 0 && xd < p.dW) dv = __ldg(p.d + ((b * p.dH + yd) * p.dW + xd) * p.Cout + o);
      }
      As[kk][cc] = av;
      Ds[kk][cc] = dv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], d[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = Ds[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], d[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = i0 + ty + 16 * i;
    if (ci >= Ctot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = o0 + tx + 16 * j;
      if (co < p.Cout) atomicAdd(p.grad + co * p.so + ci * p.si + tap * p.st, acc[i][j]);
    }
  }
}

}  // namespace gsd
