// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (SS operands) as a function of N and cta_group.
// One CTA (or CTA pair) per SM, operands are whatever is in shared memory (zero-filled); no loads in the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../divergen_b200/csrc/common.cuh"
#include "../../divergen_b200/csrc/gemm2_tc.cuh"
using namespace dg;

template <int kCta>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = (kCta == 1) || cluster_ctarank() == 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_pair<kCta, 512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && lane == 0 && leader) {
    const uint32_t idesc = make_idesc_f16((uint32_t)n, false, 128 * kCta);
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db = make_smem_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss_pair<kCta>(tmem + (i & 1) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
    }
    umma_commit_pair<kCta>(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  } else if (warp == 0 && lane == 0) {
    mbar_wait(&bar, 0);
  }
  __syncwarp();
  tc_fence_before();
  if constexpr (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair<kCta, 512>(tmem); }
}

template <int kCta>
void run(int n, int iters) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_rate_kernel<kCta>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCta == 2 ? 148 : 148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = kCta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, mma_rate_kernel<kCta>, n, iters, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cta=%d N=%d: %s\n", kCta, n, cudaGetErrorString(e)); exit(1); }
  }
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = 4.0 * iters;
  const double flops_per_sm = 2.0 * 128 * n * 16 * mmas;   // per SM (each SM of a pair computes 128 x n)
  printf("cta_group=%d N=%3d: %7.1f cycles/MMA  (%.0f MAC/clk/SM)  chip %.0f TFLOP/s at event time %.3f ms\n", kCta, n,
         (double)c / mmas, 128.0 * n * 16 * mmas / (double)c, flops_per_sm * 148 / (ms * 1e-3) / 1e12, ms);
  cudaFree(d);
}

int main() {
  const int iters = 20000;
  for (int n : {64, 96, 128, 160, 192, 224, 256}) run<1>(n, iters);
  for (int n : {64, 96, 128, 160, 192, 224, 256}) run<2>(n, iters);
  return 0;
}
