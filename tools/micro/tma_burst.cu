// Micro-benchmark: one thread per SM issues `n` TMA box loads back to back on ONE mbarrier (smem ring reused, data is
// garbage), then waits once.  Separates the TMA engine's streaming rate from per-stage synchronisation cost.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "../../divergen_b200/csrc/common.cuh"
using namespace dg;

__global__ void __launch_bounds__(128, 1) burst_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int n, int k_blocks,
                                                       int row_tiles, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int box_bytes = box_rows * 128;
    const int slots = (192 * 1024) / box_bytes;
    int rt = blockIdx.x % row_tiles, kb = (blockIdx.x * 7) % k_blocks;
    long long t0 = clock64();
    for (int rep = 0; rep < 8; ++rep) {
      mbar_arrive_expect_tx(&bar, box_bytes * n);
      for (int i = 0; i < n; ++i) {
        tma_load_2d(smem + (i % slots) * box_bytes, &map, &bar, kb * 64, rt * box_rows);
        if (++kb == k_blocks) { kb = 0; rt = (rt + gridDim.x) % row_tiles; }
      }
      mbar_wait(&bar, rep & 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int K = 4096; const size_t rows = 4096;
  void* buf; cudaMalloc(&buf, rows * K * 2); cudaMemset(buf, 0, rows * K * 2);
  for (int box_rows : {64, 128, 256}) {
    for (int n : {1, 2, 4, 8, 16, 32}) {
      if (box_rows * 128 * n > (1 << 20) - 1) continue;   // mbarrier tx-count limit
      CUtensorMap m;
      cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
      cuuint32_t bx[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
      enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (int rep = 0; rep < 2; ++rep) {
        burst_kernel<<<148, 128, 200 * 1024>>>(m, box_rows, n, K / 64, (int)(rows / box_rows), d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per_burst = (double)c / 8;
      printf("box %3d rows, %2d boxes per burst (%4d KB): %7.0f cycles/burst  %6.1f B/clk/SM\n", box_rows, n, box_rows * 128 * n / 1024,
             per_burst, (double)box_rows * 128 * n / per_burst);
    }
  }
  return 0;
}
