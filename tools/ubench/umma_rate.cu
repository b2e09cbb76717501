// Issue-rate microbenchmark of tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, K = 16) on one SM and on all SMs:
// cycles per instruction for N = 64 / 128 / 192 / 256 when R instructions are issued back to back by one elected lane,
// operands in shared memory (K-major SWIZZLE_128B tiles, contents irrelevant), into ONE accumulator or alternating
// between two.  Prints the measured cycles per UMMA = the speed of light for the conv kernels' inner loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate tools/ubench/umma_rate.cu -I gelslim_depth_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>
#include "gsd_ptx.cuh"
using namespace gsd;

template <int N, int ALT, int BG = 0>
__global__ void __launch_bounds__(128, 1) k_rate(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a = base, s_b = base + 32768, s_bar = base + 32768 + 65536, s_slot = s_bar + 16;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (s_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(s_bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(s_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = make_idesc_bf16_m128(N);
  constexpr uint32_t HI = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t a_lo = ((s_a & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((s_b & 0x3FFFFu) >> 4) | (1u << 16);
    for (int pass = 0; pass < 2; ++pass) {           // pass 0 = warm-up
      t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            umma_bf16_lohi(tmem + ((ALT && (u & 1)) ? 256 : 0), a_lo + 2 * (u & 3), HI, b_lo + 2 * (u & 3), HI, idesc, 1u);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(s_bar);
      __syncwarp();
      mbar_wait(s_bar, pass & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
    if (BG && threadIdx.x == 32) *reinterpret_cast<volatile uint32_t*>(smem_raw + (s_slot + 8 - smem_u32(smem_raw))) = 1u;
  } else if (BG && warp >= 2) {
    // background shared-memory traffic from the LSU (what an epilogue does): conflict-free 16-byte loads + stores,
    // 4 wavefronts each, until the MMA warp is done; the count goes to out[148 + block] (wavefronts)
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(smem_raw + (s_slot + 8 - smem_u32(smem_raw)));
    const uint32_t addr = base + 32768 + 65536 + 1024 + (warp - 2) * 512 + (threadIdx.x & 31) * 16;
    long long n = 0;
    uint32_t a = 1, b = 2, c = 3, d = 4;
    while (*flag == 0u) {
#pragma unroll
      for (int u = 0; u < BG; ++u) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
      }
      n += 8 * BG;
    }
    if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(out + 148 + blockIdx.x), (unsigned long long)n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Weight-stationary variant (VERDICT r1 next #1c): tcgen05.mma.ws keeps the B operand of one K step in a collector buffer
// while FOUR M tiles (four different A tiles, four accumulators) use it: fill, use, use, lastuse.  Per UMMA the shared
// memory pipe then serves A (32 wavefronts) + B/4 instead of A + B.  WS = 0 issues the same access pattern with the
// ordinary instruction (every UMMA reads its B again) for comparison.
__device__ __forceinline__ void umma_ws(uint32_t d, uint32_t a_lo, uint32_t hi, uint32_t b_lo, uint32_t idesc, int mode) {
  // mode 0 fill, 1 use, 2 lastuse
  if (mode == 0)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
  else if (mode == 1)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(hi), "r"(b_lo), "r"(idesc) : "memory");
}

template <int N, int WS>
__global__ void __launch_bounds__(128, 1) k_rate_ws(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a = base, s_b = base + 65536, s_bar = base + 65536 + 65536, s_slot = s_bar + 16;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem_raw + (s_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (65536 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(s_bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(s_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = make_idesc_bf16_m128(N);
  constexpr uint32_t HI = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t a_lo = ((s_a & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((s_b & 0x3FFFFu) >> 4) | (1u << 16);
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      for (int r = 0; r < reps; r += 16) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)            // four K steps, each shared by four M tiles
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t a = a_lo + j * (16384 >> 4) + 2 * k, b = b_lo + 2 * k, d = tmem + j * N;
              if (WS) umma_ws(d, a, HI, b, idesc, j == 0 ? 0 : (j == 3 ? 2 : 1));
              else umma_bf16_lohi(d, a, HI, b, HI, idesc, 1u);
            }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(s_bar);
      __syncwarp();
      mbar_wait(s_bar, pass & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N, int WS>
void run_ws(int grid, int reps, long long* d_out) {
  const int smem = 65536 + 65536 + 4096;
  cudaMemset(d_out, 0, sizeof(long long) * 296);
  cudaFuncSetAttribute(k_rate_ws<N, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_rate_ws<N, WS><<<grid, 128, smem>>>(reps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ws N=%d ws=%d grid=%d: %s\n", N, WS, grid, cudaGetErrorString(e)); return; }
  long long h[296];
  cudaMemcpy(h, d_out, sizeof(long long) * 296, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d four M tiles per K step, %s, SMs=%3d: %.1f cycles per UMMA (nominal %d)  -> %.0f FLOP/clk/SM\n", N,
         WS ? "tcgen05.mma.ws collector::b0 fill/use/use/lastuse" : "ordinary tcgen05.mma", grid, (double)mx / reps, N / 2,
         2.0 * 128 * N * 16 * reps / (double)mx);
}

template <int N, int ALT, int BG = 0>
void run(int grid, int reps, long long* d_out) {
  const int smem = 32768 + 65536 + 4096;
  cudaMemset(d_out, 0, sizeof(long long) * 296);
  cudaFuncSetAttribute(k_rate<N, ALT, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_rate<N, ALT, BG><<<grid, 128, smem>>>(reps, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d alt=%d grid=%d: %s\n", N, ALT, grid, cudaGetErrorString(e)); return; }
  long long h[296];
  cudaMemcpy(h, d_out, sizeof(long long) * 296, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d accumulators=%d SMs=%3d bg=%d: %.1f cycles per UMMA (nominal %d)  -> %.0f FLOP/clk/SM", N, ALT ? 2 : 1, grid, BG,
         (double)mx / reps, N / 2, 2.0 * 128 * N * 16 * reps / (double)mx);
  if (BG) printf("   LSU smem wavefronts per UMMA: %.1f (per cycle %.2f)", (double)h[148] / reps / 2, (double)h[148] / 2 / (double)h[0]);
  printf("\n");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 296);
  const int reps = 4096;
  for (int grid : {1, 148}) {
    run<64, 0>(grid, reps, d_out);  run<64, 1>(grid, reps, d_out);
    run<128, 0>(grid, reps, d_out); run<128, 1>(grid, reps, d_out);
    run<192, 0>(grid, reps, d_out);
    run<256, 0>(grid, reps, d_out); run<256, 1>(grid, reps, d_out);
  }
  // shared-memory pipe contention: the same UMMA streams with two LSU warps hammering shared memory
  run<64, 0, 1>(1, reps, d_out); run<128, 0, 1>(1, reps, d_out); run<256, 0, 1>(1, reps, d_out);
  run<128, 0, 4>(1, reps, d_out); run<256, 0, 4>(1, reps, d_out);
  // weight-stationary B (collector) across four M tiles vs the ordinary instruction on the same access pattern
  for (int grid : {1, 148}) {
    run_ws<64, 0>(grid, reps, d_out);  run_ws<64, 1>(grid, reps, d_out);
    run_ws<128, 0>(grid, reps, d_out); run_ws<128, 1>(grid, reps, d_out);
  }
  return 0;
}
