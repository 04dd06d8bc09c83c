// Micro-benchmark: MUFU.EX2 throughput, fp32 vs packed f16x2 (elements per clock per SM), 8 warps x 4 SMSPs resident.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

template <int kMode>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8]; uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + i; }   // ~ -0.5 in both halves
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kMode == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + (float)h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 8);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(o, iters, c); else k<1><<<148, threads>>>(o, iters, c);
        cudaDeviceSynchronize();
      }
      long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
      const double instr = (double)iters * 8 * threads;     // thread-level ex2 instructions per SM
      printf("%s threads/SM %4d: %.2f thread-instr/clk/SM = %.1f elements/clk/SM\n", mode ? "ex2.f16x2" : "ex2.f32  ", threads,
             instr / cy, instr / cy * (mode ? 2 : 1));
    }
  }
  return 0;
}
