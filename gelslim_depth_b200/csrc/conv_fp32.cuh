// fp32 parity path: the same network as the bf16 tensor-core path, computed with plain FFMA implicit-GEMM
// kernels on fp32 NHWC activations.  tcgen05 has no true-fp32 MMA (kind::tf32 keeps 10 mantissa bits), so the
// "max-abs depth error <= 1e-3 mm in fp32" bound of the north star is met by this CUDA-core mode; it is a
// PARITY mode (about 1/40 of the bf16 path's speed), selected with dtype = GSD_DTYPE_FP32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsd {

struct F32Conv {
  const float* src0; int C0;                      // (B,H,W,C0)
  const float* src1; int C1, H1, W1, off_y, off_x; // (B,H1,W1,C1) placed at (off_y, off_x): virtual pad + concat
  const float* w;                                 // [ntaps][C0+C1][Ntot], Ntot contiguous
  const float* scale; const float* shift;         // [Ntot]
  float* out;                                     // (B,H,W,Cout) or (B,2H,2W,Cout) when groups == 4
  int B, H, W, Cout, groups, ntaps, relu;
  int8_t dy[9], dx[9];
  // training-path extensions (0 = the inference defaults): leading dimension of w when only a column range of it is used,
  // and a strided / offset sampling of src0: sample (in_stride * y + dy + s_oy, in_stride * x + dx + s_ox) of a (B, sH, sW, C0) tensor
  int ldw = 0, in_stride = 0, sH = 0, sW = 0, s_oy = 0, s_ox = 0;
};

// 64 pixels x 64 channels per block, 256 threads, 4x4 outputs per thread, K walked in chunks of 16.
__global__ void __launch_bounds__(256) conv_f32_kernel(const F32Conv p) {
  __shared__ float As[16][65];   // [k][pixel]
  __shared__ float Bs[16][64];   // [k][channel]
  const int Ctot = p.C0 + p.C1;
  const int K = p.ntaps * Ctot;
  const int Ntot = p.groups * p.Cout;
  const long M = (long)p.B * p.H * p.W;
  const long m0 = (long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int ldw = p.ldw ? p.ldw : Ntot;
  const int S = p.in_stride ? p.in_stride : 1, sH = p.sH ? p.sH : p.H, sW = p.sW ? p.sW : p.W;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx -> channels, ty -> pixels
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 16) {
    // A chunk: 64 pixels x 16 k; thread -> (kk = t % 16, pixel = t / 16 + 16*r)
    {
      const int kk = threadIdx.x & 15;
      const int k = k0 + kk;
      int tap = 0, c = 0;
      if (k < K) { tap = k / Ctot; c = k - tap * Ctot; }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int pl = (threadIdx.x >> 4) + 16 * r;
        const long m = m0 + pl;
        float v = 0.f;
        if (k < K && m < M) {
          const int x = (int)(m % p.W);
          const int y = (int)((m / p.W) % p.H);
          const long b = m / ((long)p.W * p.H);
          const int ys = y + p.dy[tap], xs = x + p.dx[tap];
          if (c < p.C0) {
            const int y0 = S * y + p.dy[tap] + p.s_oy, x0 = S * x + p.dx[tap] + p.s_ox;
            if (y0 >= 0 && y0 < sH && x0 >= 0 && x0 < sW) v = __ldg(p.src0 + ((b * sH + y0) * sW + x0) * p.C0 + c);
          } else {
            const int y1 = ys - p.off_y, x1 = xs - p.off_x;
            if (y1 >= 0 && y1 < p.H1 && x1 >= 0 && x1 < p.W1)
              v = __ldg(p.src1 + ((b * p.H1 + y1) * p.W1 + x1) * p.C1 + (c - p.C0));
          }
        }
        As[kk][pl] = v;
      }
    }
    // B chunk: 16 k x 64 channels
    {
      const int nn = threadIdx.x & 63;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int kk = (threadIdx.x >> 6) + 4 * r;
        const int k = k0 + kk;
        Bs[kk][nn] = (k < K && n0 + nn < Ntot) ? __ldg(p.w + (long)k * ldw + n0 + nn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long m = m0 + ty + 16 * i;
    if (m >= M) continue;
    const int x = (int)(m % p.W);
    const int y = (int)((m / p.W) % p.H);
    const long b = m / ((long)p.W * p.H);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= Ntot) continue;
      float v = acc[i][j];
      if (p.scale && p.shift) v = v * p.scale[n] + p.shift[n];
      else if (p.scale) v *= p.scale[n];
      else if (p.shift) v += p.shift[n];
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.groups == 1) {
        p.out[((b * p.H + y) * p.W + x) * p.Cout + n] = v;
      } else {
        const int g = n / p.Cout, co = n - g * p.Cout;
        p.out[((b * 2 * p.H + 2 * y + (g >> 1)) * (2L * p.W) + 2 * x + (g & 1)) * p.Cout + co] = v;
      }
    }
  }
}

// MaxPool2d(2), floor mode (unet.py:26), fp32 NHWC
__global__ void __launch_bounds__(256) maxpool_f32_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                                          float* __restrict__ out) {
  const int Hp = H / 2, Wp = W / 2;
  const long total = (long)B * Hp * Wp * C;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int x = (int)((idx / C) % Wp);
    const int y = (int)((idx / ((long)C * Wp)) % Hp);
    const long b = idx / ((long)C * Wp * Hp);
    const float* s = in + ((b * H + 2 * y) * W + 2 * x) * C + c;
    out[idx] = fmaxf(fmaxf(s[0], s[C]), fmaxf(s[(long)W * C], s[(long)W * C + C]));
  }
}

// OutConv 1x1 + bias + de-normalisation on fp32 NHWC -> fp32 NCHW
__global__ void __launch_bounds__(256) head_f32_kernel(const float* __restrict__ in, int Cin, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int ncls, float out_scale,
                                                       float out_shift, long npix_per_img, int B, float* __restrict__ y) {
  const long total = npix_per_img * B * ncls;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long pix = idx % npix_per_img;
    const int k = (int)((idx / npix_per_img) % ncls);
    const long b = idx / (npix_per_img * ncls);
    const float* s = in + (b * npix_per_img + pix) * Cin;
    float acc = 0.f;
    for (int c = 0; c < Cin; ++c) acc = fmaf(s[c], w[k * Cin + c], acc);
    y[idx] = (acc + bias[k]) * out_scale + out_shift;
  }
}

// Conv2d weight (O, I, kh, kw) -> [tap][I][O]
__global__ void pack_conv_weight_f32_kernel(const float* __restrict__ w, int O, int I, int taps, float* __restrict__ out) {
  const long total = (long)O * I * taps;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % O);
    const int i = (int)((idx / O) % I);
    const int t = (int)(idx / ((long)O * I));
    out[idx] = w[((long)o * I + i) * taps + t];
  }
}
// ConvTranspose2d weight (I, O, 2, 2) -> [I][4*O] with n = (dy*2+dx)*O + o
__global__ void pack_convt_weight_f32_kernel(const float* __restrict__ w, int I, int O, float* __restrict__ out) {
  const long total = 4L * O * I;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % (4 * O));
    const int i = (int)(idx / (4 * O));
    const int g = n / O, o = n - g * O;
    out[idx] = w[((long)i * O + o) * 4 + g];
  }
}

}  // namespace gsd
