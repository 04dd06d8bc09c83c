// Micro-benchmark: sustained TMA tile-load throughput per SM (all SMs active), GEMM-like access pattern.
// Every CTA streams [rows x 64]-element boxes (128-byte swizzle) of a row-major fp16 matrix [R, K] through a ring of
// `stages` shared-memory buffers; nothing consumes the data.  Prints bytes/clk/SM and TB/s for an L2-resident and an
// HBM-streaming footprint.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "../../divergen_b200/csrc/common.cuh"
using namespace dg;

__global__ void __launch_bounds__(128, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int boxes_per_stage,
                                                          int stages, int iters, int k_blocks, int row_tiles, long long* cycles, int nw, int wait_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_all[64];
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&full_all[i], 1); fence_barrier_init(); }
  __syncthreads();
  const int wi = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && wi < nw) {
    uint64_t* full = full_all + wi * 16;
    const int stage_bytes = box_rows * 128 * boxes_per_stage;
    smem += wi * stages * stage_bytes;
    long long t0 = clock64();
    int issued = 0, done = 0;
    // GEMM-like walk: this CTA owns a row tile and streams along K, then moves on
    int rt = (blockIdx.x * 4 + wi) % row_tiles, kb = (blockIdx.x * 7) % k_blocks;
    while (done < iters) {
      while (issued < iters && issued - done < stages) {
        const int s = issued % stages;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        for (int b = 0; b < boxes_per_stage; ++b)
          tma_load_2d(smem + s * stage_bytes + b * box_rows * 128, &map, &full[s], kb * 64, ((rt * boxes_per_stage + b) * box_rows));
        if (++kb == k_blocks) { kb = 0; rt = (rt + gridDim.x) % row_tiles; }
        ++issued;
      }
      if (wait_mode == 0) mbar_wait(&full[done % stages], (done / stages) & 1);
      else {
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&full[done % stages])), "r"((uint32_t)((done / stages) & 1)) : "memory");
        }
      }
      ++done;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && wi == 0) *cycles = t1 - t0;
  }
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  void* p = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int K = 4096;
  struct Cfg { int box_rows, bps, stages, nw, wm; };
  const Cfg cfgs[] = {{128, 1, 4, 1, 0}, {128, 1, 4, 1, 1}, {256, 1, 4, 1, 1}, {128, 2, 4, 1, 1}, {80, 2, 4, 1, 1}, {128, 1, 4, 2, 1},
                      {128, 1, 2, 4, 1}, {64, 1, 4, 1, 1}, {32, 1, 4, 1, 1}, {128, 1, 1, 1, 1}, {128, 1, 1, 1, 0}};
  for (size_t rows : {(size_t)4096}) {
    void* buf; cudaMalloc(&buf, rows * K * 2); cudaMemset(buf, 0, rows * K * 2);
    for (const Cfg& c_ : cfgs) {
      const int box_rows = c_.box_rows, bps = c_.bps, stages = c_.stages, nw = c_.nw, wm = c_.wm;
      CUtensorMap m;
      cuuint64_t gd[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t gs[1] = {(cuuint64_t)K * 2};
      cuuint32_t bx[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
      enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const int iters = 4000;
      const int row_tiles = (int)(rows / (box_rows * bps));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        tma_rate_kernel<<<148, 128, 200 * 1024>>>(m, box_rows, bps, stages, iters, K / 64, row_tiles, d, nw, wm);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * box_rows * 128 * bps * nw;
      printf("wait %d box %3d rows x%d  stages %2d  warps %d (%3d KB in flight): %6.1f B/clk/SM  %6.0f cycles/stage/warp\n",
             wm, box_rows, bps, stages, nw, box_rows * 128 * bps * stages * nw / 1024, bytes / (double)c, (double)c / iters);
    }
    cudaFree(buf);
  }
  return 0;
}
