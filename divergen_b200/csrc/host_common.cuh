// Host-side plumbing shared by the C-ABI translation unit: error reporting, TMA tensor-map construction through the
// driver entry point (no link-time libcuda dependency), kernel launchers for the tcgen05 GEMM/conv and attention
// kernels and for the HBM-bound elementwise kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/divergen_b200.h"
#include "attn_tc.cuh"
#include "elementwise.cuh"
#include "gemm_tc.cuh"

namespace dg {

inline thread_local std::string g_last_error;
inline thread_local long long g_launch_counter = 0;

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define DG_CUDA(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess) return ::dg::fail(DG_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                             __FILE__, __LINE__);                                          \
  } while (0)
#define DG_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != DG_OK) return _r; \
  } while (0)
#define DG_LAUNCH_CHECK()                                                                                   \
  do {                                                                                                      \
    ++::dg::g_launch_counter;                                                                               \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) return ::dg::fail(DG_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                             __FILE__, __LINE__);                                           \
  } while (0)

// ------------------------------------------------------------------ per-launch profiler (bench.py roofline numbers)
// When active, every launcher below brackets its kernel with a CUDA-event pair on the launch stream and records the
// algorithmic FLOPs / bytes of that launch.  Never active during graph capture or the timed bench region.
enum Family { FAM_GEMM = 0, FAM_ATTN = 1, FAM_NORM = 2, FAM_OTHER = 3, FAM_COUNT = 4 };
struct Profiler {
  struct Rec { int fam; cudaEvent_t a, b; double flops, bytes; };
  std::vector<Rec> recs;
};
inline thread_local Profiler* g_prof = nullptr;
struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(int fam, cudaStream_t s_, double flops, double bytes) : s(s_), on(g_prof != nullptr) {
    if (!on) return;
    Profiler::Rec r{fam, nullptr, nullptr, flops, bytes};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    g_prof->recs.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(g_prof->recs.back().b, s); }
};

// DG_TRACE=1: one stderr line per tcgen05 launch (shape census that lines up with an ncu launch list).
inline bool trace_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DG_TRACE"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// ------------------------------------------------------------------ tensor maps
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// rank-4 fp16 map, 128-byte swizzle, zero OOB fill.  dims/box innermost first; strides (bytes) for dims 1..3.
inline int make_map_4d(CUtensorMap* m, const void* ptr, const uint64_t dims[4], const uint64_t strides[3],
                       const uint32_t box[4], bool weights = false, bool swizzle64 = false) {
  auto fn = get_encode_fn();
  if (!fn) return fail(DG_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  weights ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DG_E_CUDA,
                "cuTensorMapEncodeTiled(4d) failed: %d ptr=%p dims={%llu,%llu,%llu,%llu} strides={%llu,%llu,%llu} "
                "box={%u,%u,%u,%u}",
                (int)r, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
                (unsigned long long)strides[2], box[0], box[1], box[2], box[3]);
  return DG_OK;
}
inline int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride_bytes,
                       uint32_t box_inner, uint32_t box_outer, bool weights = true) {
  auto fn = get_encode_fn();
  if (!fn) return fail(DG_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[2] = {inner, outer};
  cuuint64_t gs[1] = {stride_bytes};
  cuuint32_t bx[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  weights ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DG_E_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d ptr=%p dims={%llu,%llu} stride=%llu box={%u,%u}",
                (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride_bytes,
                box_inner, box_outer);
  return DG_OK;
}

// ------------------------------------------------------------------ GEMM / conv launcher
constexpr int kGemmBlockN = 160;   // divides every channel count of SD-1.x/2.x (320, 640, 1280, ... 10240)
constexpr int kGemmStages = 5;     // 5 x 36 KB ring + 40 KB staging tile
constexpr int kGegluBlockN = 128;  // GEGLU tiles: [64 value | 64 gate] -> 64 outputs
constexpr int kGegluStages = 6;

struct GemmArgs {
  const __half* a0 = nullptr; int c0 = 0;  // source 0: NHWC [B,H,W,c0]
  const __half* a1 = nullptr; int c1 = 0;  // optional source 1 (channel concat)
  int B = 1, H = 1, W = 1;                 // plain GEMM: B = H = 1, W = M
  int taps = 1;                            // 1 (Linear / 1x1) or 9 (3x3, stride 1, pad 1)
  const __half* w = nullptr;               // packed [n_w, taps*(c0+c1)]
  int n_w = 0;                             // rows of w
  int n_out = 0;                           // output columns
  const __half* bias = nullptr;
  const __half* rowvec = nullptr; int ld_rowvec = 0;
  const __half* residual = nullptr; int ld_res = 0;
  int geglu = 0;
  __half* out = nullptr; int ldo = 0;
};

inline int largest_pow2_divisor(int x, int cap) {
  int p = 1;
  while (p * 2 <= cap && x % (p * 2) == 0) p *= 2;
  return p;
}

inline int launch_gemm(cudaStream_t stream, int num_sms, const GemmArgs& a) {
  if (a.c0 % 64 || a.c1 % 64 || a.c0 <= 0) return fail(DG_E_SHAPE, "gemm: channel counts must be multiples of 64 (%d,%d)", a.c0, a.c1);
  if (a.taps != 1 && a.taps != 9) return fail(DG_E_ARG, "gemm: taps must be 1 or 9");
  if ((reinterpret_cast<uintptr_t>(a.a0) | reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.out) |
       reinterpret_cast<uintptr_t>(a.residual) | reinterpret_cast<uintptr_t>(a.a1)) & 15)
    return fail(DG_E_ARG, "gemm: pointers must be 16-byte aligned");
  if (a.ldo % 8 || (a.residual && a.ld_res % 8)) return fail(DG_E_SHAPE, "gemm: output / residual row pitch must be a multiple of 8 elements");
  const int block_n = a.geglu ? kGegluBlockN : kGemmBlockN;
  GemmParams p{};
  int W = a.W, H = a.H, B = a.B;
  if (a.taps == 1) { W = a.B * a.H * a.W; H = 1; B = 1; }
  p.W = W; p.H = H; p.B = B;
  p.bw = largest_pow2_divisor(W, 128);
  if (a.taps == 1) p.bw = 128;
  p.bh = (a.taps == 1) ? 1 : largest_pow2_divisor(H, 128 / p.bw);
  p.bn = 128 / (p.bw * p.bh);
  p.tiles_x = (W + p.bw - 1) / p.bw;
  p.tiles_y = (H + p.bh - 1) / p.bh;
  p.tiles_b = (B + p.bn - 1) / p.bn;
  p.n_gemm = a.n_w;
  p.tiles_n = (a.n_w + block_n - 1) / block_n;
  p.n_out = a.n_out;
  p.taps = a.taps;
  p.kb0 = a.c0 / 64; p.kb1 = a.c1 / 64;
  p.bias = a.bias; p.rowvec = a.rowvec; p.ld_rowvec = a.ld_rowvec;
  p.has_residual = a.residual != nullptr;
  if (a.geglu && !a.bias) return fail(DG_E_ARG, "gemm: geglu epilogue needs a packed bias");

  CUtensorMap mA0, mA1, mW, mO, mR;
  {
    uint64_t dims[4] = {(uint64_t)a.c0, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t st[3] = {(uint64_t)a.c0 * 2, (uint64_t)W * a.c0 * 2, (uint64_t)H * W * a.c0 * 2};
    uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    DG_TRY(make_map_4d(&mA0, a.a0, dims, st, box));
    if (a.c1 > 0) {
      uint64_t d1[4] = {(uint64_t)a.c1, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      uint64_t s1[3] = {(uint64_t)a.c1 * 2, (uint64_t)W * a.c1 * 2, (uint64_t)H * W * a.c1 * 2};
      DG_TRY(make_map_4d(&mA1, a.a1, d1, s1, box));
    } else {
      mA1 = mA0;
    }
    const uint64_t ktot = (uint64_t)a.taps * (a.c0 + a.c1);
    DG_TRY(make_map_2d(&mW, a.w, ktot, (uint64_t)a.n_w, ktot * 2, 64, block_n));
    // output / residual tiles: {n_out, W, H, B} boxes of 32 columns x 128 pixels, 64-byte swizzle in smem
    uint64_t od[4] = {(uint64_t)a.n_out, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t os[3] = {(uint64_t)a.ldo * 2, (uint64_t)W * a.ldo * 2, (uint64_t)H * W * a.ldo * 2};
    uint32_t obox[4] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    DG_TRY(make_map_4d(&mO, a.out, od, os, obox, false, true));
    if (a.residual) {
      uint64_t rs[3] = {(uint64_t)a.ld_res * 2, (uint64_t)W * a.ld_res * 2, (uint64_t)H * W * a.ld_res * 2};
      DG_TRY(make_map_4d(&mR, a.residual, od, rs, obox, false, true));
    } else {
      mR = mO;
    }
  }
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.tiles_n;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (trace_on())
    fprintf(stderr, "DG_TRACE gemm M=%d N=%d K=%d taps=%d c0=%d c1=%d tiles=%d grid=%d geglu=%d\n", a.B * a.H * a.W, a.n_w,
            a.taps * (a.c0 + a.c1), a.taps, a.c0, a.c1, total_tiles, grid, a.geglu);
  const double rows_ = (double)a.B * a.H * a.W, ktot_ = (double)a.taps * (a.c0 + a.c1);
  ProfScope prof_(FAM_GEMM, stream, 2.0 * rows_ * ktot_ * (double)a.n_w,
                  2.0 * (rows_ * (a.c0 + a.c1) + ktot_ * a.n_w + rows_ * a.n_out));
  if (a.geglu)
    gemm_tc_kernel<kGegluBlockN, kGegluStages, true>
        <<<grid, 384, GemmSmem<kGegluBlockN, kGegluStages, true>::kTotal, stream>>>(mA0, mA1, mW, mO, mR, p);
  else
    gemm_tc_kernel<kGemmBlockN, kGemmStages, false>
        <<<grid, 384, GemmSmem<kGemmBlockN, kGemmStages, false>::kTotal, stream>>>(mA0, mA1, mW, mO, mR, p);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

// ------------------------------------------------------------------ attention launcher
template <int kD, int kKV, int kStages>
inline int launch_attn_t(cudaStream_t stream, const __half* q, int ldq, const __half* k, int ldk, const __half* v,
                         int ldv, __half* out, int B, int heads, int Sq, int Sk) {
  using C = AttnCfg<kD, kKV, kStages>;
  CUtensorMap mQ, mK, mV;
  auto mk = [&](CUtensorMap* m, const __half* ptr, int ld, int S, int rows) -> int {
    uint64_t dims[4] = {(uint64_t)kD, (uint64_t)heads, (uint64_t)S, (uint64_t)B};
    uint64_t st[3] = {(uint64_t)kD * 2, (uint64_t)ld * 2, (uint64_t)S * ld * 2};
    uint32_t box[4] = {64, 1, (uint32_t)rows, 1};
    return make_map_4d(m, ptr, dims, st, box);
  };
  DG_TRY(mk(&mQ, q, ldq, Sq, 128));
  DG_TRY(mk(&mK, k, ldk, Sk, kKV));
  DG_TRY(mk(&mV, v, ldv, Sk, kKV));
  AttnParams p{};
  p.Sq = Sq; p.Sk = Sk; p.heads = heads; p.ldo = heads * kD; p.out = out;
  p.scale_log2 = (1.0f / sqrtf((float)kD)) * 1.4426950408889634f;
  auto kern = attn_tc_kernel<kD, kKV, kStages>;
  dim3 grid((Sq + 255) / 256, heads, B);
  if (trace_on()) fprintf(stderr, "DG_TRACE attn B=%d heads=%d Sq=%d Sk=%d d=%d\n", B, heads, Sq, Sk, kD);
  ProfScope prof_(FAM_ATTN, stream, 4.0 * B * heads * (double)Sq * Sk * kD,
                  2.0 * B * heads * kD * (2.0 * Sq + 2.0 * Sk));
  kern<<<grid, 384, C::kSmem, stream>>>(mQ, mK, mV, p);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

inline int launch_attention(cudaStream_t stream, const __half* q, int ldq, const __half* k, int ldk, const __half* v,
                            int ldv, __half* out, int B, int heads, int Sq, int Sk, int d) {
  if ((ldq | ldk | ldv) % 8) return fail(DG_E_SHAPE, "attention: row strides must be multiples of 8 elements");
  switch (d) {
    case 32: return launch_attn_t<32, 128, 4>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    case 40: return launch_attn_t<40, 128, 4>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    case 64: return launch_attn_t<64, 128, 3>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    case 80: return launch_attn_t<80, 128, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    case 160: return launch_attn_t<160, 64, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    default: return fail(DG_E_UNSUPPORTED, "attention: head dim %d not built (32/40/64/80/160)", d);
  }
}

// Opt every tcgen05 kernel into its dynamic shared-memory size once per device (never during graph capture).
template <int kD, int kKV, int kStages>
inline int init_attn_attr() {
  DG_CUDA(cudaFuncSetAttribute(attn_tc_kernel<kD, kKV, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               AttnCfg<kD, kKV, kStages>::kSmem));
  return DG_OK;
}
inline int init_kernel_attributes() {
  DG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kGemmBlockN, kGemmStages, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmSmem<kGemmBlockN, kGemmStages, false>::kTotal));
  DG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kGegluBlockN, kGegluStages, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmSmem<kGegluBlockN, kGegluStages, true>::kTotal));
  DG_TRY((init_attn_attr<32, 128, 4>()));
  DG_TRY((init_attn_attr<40, 128, 4>()));
  DG_TRY((init_attn_attr<64, 128, 3>()));
  DG_TRY((init_attn_attr<80, 128, 2>()));
  DG_TRY((init_attn_attr<160, 64, 2>()));
  return DG_OK;
}

// ------------------------------------------------------------------ elementwise launchers
inline int grid_for(size_t work_items, int block, int num_sms, int waves = 8) {
  size_t g = (work_items + block - 1) / block;
  size_t cap = (size_t)num_sms * waves;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

inline int launch_groupnorm(cudaStream_t s, int num_sms, const __half* x0, int C0, const __half* x1, int C1,
                            const __half* gamma, const __half* beta, __half* out, float* stats, int B, int HW,
                            int groups, float eps, int silu) {
  const int C = C0 + C1;
  if (C % groups || C0 % 8 || C1 % 8) return fail(DG_E_SHAPE, "groupnorm: C=%d+%d groups=%d", C0, C1, groups);
  ProfScope prof_(FAM_NORM, s, 0.0, 2.0 * 3.0 * B * HW * (double)C);
  DG_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * B, s));
  const int nvec = C / 8;
  const int pstride = nvec <= 256 ? 256 / nvec : 1;
  // strip length: enough blocks to fill the chip (>= 4 per SM when the tensor allows), >= 8 pixel iterations per thread
  int ppb = 8 * pstride;
  while (ppb * 2 <= HW && (size_t)B * ((HW + ppb - 1) / ppb) > (size_t)8 * num_sms) ppb *= 2;
  dim3 grid((HW + ppb - 1) / ppb, B);
  const size_t smem = sizeof(float) * 2 * (size_t)pstride * C;
  gn_stats_kernel<<<grid, 256, smem, s>>>(x0, C0, x1, C1, HW, groups, ppb, stats);
  DG_LAUNCH_CHECK();
  gn_apply_kernel<<<grid, 256, 0, s>>>(x0, C0, x1, C1, HW, groups, eps, ppb, stats, gamma, beta, silu, out);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

inline int launch_layernorm(cudaStream_t s, const __half* x, const __half* gamma, const __half* beta, __half* out,
                            int rows, int C, float eps) {
  if (C % 8 || C > 5 * 256) return fail(DG_E_SHAPE, "layernorm: C=%d unsupported", C);
  const int warps = 8;
  ProfScope prof_(FAM_NORM, s, 0.0, 2.0 * 2.0 * rows * (double)C);
  layernorm_kernel<5><<<(rows + warps - 1) / warps, warps * 32, 0, s>>>(x, gamma, beta, out, rows, C, eps);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

inline int launch_gemv(cudaStream_t s, const __half* x, int ldx, const __half* W, const __half* bias, __half* out,
                       int ldo, int B, int N, int K, int silu_in, int silu_out) {
  if (B > 8 || K % 8) return fail(DG_E_SHAPE, "gemv: B=%d K=%d unsupported", B, K);
  const int warps = 8, rows_per_block = warps * kGemvRowsPerWarp;
  const size_t smem = sizeof(float) * B * K;
  ProfScope prof_(FAM_OTHER, s, 2.0 * B * (double)N * K, 2.0 * ((double)N * K + B * (double)(N + K)));
  gemv_small_batch_kernel<<<(N + rows_per_block - 1) / rows_per_block, warps * 32, smem, s>>>(x, ldx, W, bias, out, ldo, B, N, K, silu_in,
                                                                           silu_out);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

}  // namespace dg
