// Host-side plumbing shared by the C-ABI translation unit: error reporting, TMA tensor-map construction through the
// driver entry point (no link-time libcuda dependency), kernel launchers for the tcgen05 GEMM/conv and attention
// kernels and for the HBM-bound elementwise kernels.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/divergen_b200.h"
#include "attn_tc.cuh"
#include "elementwise.cuh"
#include "gemm2_tc.cuh"

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

// ------------------------------------------------------------------ launches
// DG_PDL=0 disables programmatic dependent launch (A/B runs).
inline bool pdl_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DG_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// Launch a kernel whose first global access is behind griddep_wait() so that its prologue overlaps the predecessor's tail.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_on()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
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
                       const uint32_t box[4], bool weights = false, bool swizzle64 = false, uint32_t pixel_stride = 1) {
  (void)weights;
  auto fn = get_encode_fn();
  if (!fn) return fail(DG_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides[0], strides[1], strides[2]};
  // pixel_stride 2: the box takes every second pixel along W and H (boxDim counts the span in the tensor, the tile in shared
  // memory is box[1] x box[2] pixels) -- the output map of one phase of an upsample convolution
  cuuint32_t bx[4] = {box[0], box[1] * pixel_stride, box[2] * pixel_stride, box[3]};
  cuuint32_t es[4] = {1, pixel_stride, pixel_stride, 1};
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
constexpr int kGemmTileN = 320;     // widest tile: two 160-wide accumulators
// GEGLU packing: tiles of kGegluTile weight rows = [kGegluHalf value | kGegluHalf gate | zero rows].  Default 320/160 on
// the 320-wide kernel.  DG_NVCC_EXTRA=-DDG_GEGLU_NARROW builds 160/64 on the double-buffered 160-wide kernel (epilogue
// overlaps the next tile's MMA) -- measured SLOWER on the whole forward (10.20 vs 9.95 ms): 2.5x more tiles, each paying
// the ~1.2 k-cycle gap between tiles, and a fifth of the MMA columns wasted.
// Round 2: 256/128 -- two 128-wide accumulators, DOUBLE-buffered in TMEM (2 x 256 of 512 columns) with two epilogue sets: the
// packed-GELU epilogue (~2x the MMA time of a K = 320 tile) overlaps the next tile's main loop, and two tiles' epilogues run
// at once.  DG_NVCC_EXTRA=-DDG_GEGLU_WIDE builds round 1's 320/160 single-stage layout for A/B runs.
#if defined(DG_GEGLU_NARROW)
constexpr int kGegluTile = 160, kGegluHalf = 64;
#elif defined(DG_GEGLU_WIDE)
constexpr int kGegluTile = 320, kGegluHalf = 160;
#else
constexpr int kGegluTile = 256, kGegluHalf = 128;
#endif
constexpr int kGegluSets = kGegluTile == 320 ? 1 : 2;
// smem ring depths per (CTAs per tile, tile N): stage bytes are 16 KB of A + 10/20/40 KB of B; + 64 KB output staging ring
constexpr int kStages_2_320 = 4;    // 4 x 36 KB
constexpr int kStages_2_160 = 5;    // 5 x 26 KB (+ 80 KB: two whole-tile staging slots)
constexpr int kStages_1_320 = 2;    // 2 x 56 KB (single-CTA variants: bring-up / A-B runs only, DG_GEMM_CTA=1)
constexpr int kStages_1_160 = 3;    // 3 x 36 KB
constexpr int kStages_2_256 = 4;    // 4 x 32 KB (+ 64 KB: two whole-tile staging slots of 128 output columns)
constexpr int kStages_1_256 = 3;    // 3 x 48 KB
constexpr int kStagesGeglu2 = kGegluTile == 320 ? kStages_2_320 : kGegluTile == 256 ? kStages_2_256 : kStages_2_160;
constexpr int kStagesGeglu1 = kGegluTile == 320 ? kStages_1_320 : kGegluTile == 256 ? kStages_1_256 : kStages_1_160;
constexpr size_t kSplitWsFloats = (size_t)8 << 20;    // 32 MB of fp32 split-K tile accumulators (zero between launches)
constexpr int kSplitTickets = 1 << 16;

// Device resources shared by every GEMM launch of a context.
struct GemmRes {
  int num_sms = 0;
  int max_pairs = 0;        // co-resident 2-CTA clusters of the pair kernel (cudaOccupancyMaxActiveClusters)
  int cta_mode = 2;         // 2 = CTA pairs (default), 1 = single-CTA tiles (DG_GEMM_CTA=1, bring-up / A-B runs)
  float* ws = nullptr;      // split-K partials
  int* tickets = nullptr;   // split-K arrival counters (self-resetting)
};

struct GemmArgs {
  const __half* a0 = nullptr; int c0 = 0;  // source 0: NHWC [B,H,W,c0]
  const __half* a1 = nullptr; int c1 = 0;  // optional source 1 (channel concat)
  const __half* x0 = nullptr; int cx0 = 0; // extra 1x1 sources (same B, H, W): their channels follow the taps in K -- the weight is
  const __half* x1 = nullptr; int cx1 = 0; //   [n_w, taps*(c0+c1) + cx0 + cx1]; ResnetBlock2D's conv_shortcut accumulated inside conv2
  int B = 1, H = 1, W = 1;                 // plain GEMM: B = H = 1, W = M
  int taps = 1;                            // 1 (Linear / 1x1), 9 (3x3, stride 1, pad 1) or 4 (one 2x2 phase of nearest-x2 + conv3x3)
  int in_stride = 1;                       // taps == 9: 2 = stride-2 conv (Downsample2D): B/H/W are the OUTPUT grid, the source is [B, 2H, 2W, c0]
  int phase_x = 0, phase_y = 0;            // taps == 4: output pixels (2x + phase_x, 2y + phase_y) of a [B, 2H, 2W, n_out] tensor
  int gn_slot0 = 0, gn_slots = 0;          // taps == 4: this phase's first GroupNorm slab / slabs per sample of the whole output
  int hw = 0;                              // plain GEMM: rows per sample (needed for gn_stats)
  const __half* w = nullptr;               // packed [n_w, taps*(c0+c1)]
  int n_w = 0;                             // rows of w
  int n_out = 0;                           // output columns
  const __half* bias = nullptr;
  const float* bias32 = nullptr;           // fp32 bias (LayerNorm-folded layers)
  const float* colsum = nullptr;           // LayerNorm fold
  const float* ln_stats = nullptr; int ln_parts = 0; int ln_c = 0; float ln_eps = 1e-5f;
  const __half* rowvec = nullptr; int ld_rowvec = 0;
  const __half* residual = nullptr; int ld_res = 0;
  int geglu = 0;
  float* row_stats_out = nullptr;          // [rows][2*tiles_n][2]
  float* gn_stats_out = nullptr; int gn_blk = 0;   // [B][HW/32][n_out/gn_blk][2] (see gn_stats_supported)
  __half* out = nullptr; int ldo = 0;
  // fused GroupNorm (+ SiLU) on the A operand: [B][3][c0 + c1] fp16 planes from launch_gn_fold (nullptr: A is used as is)
  const __half* xf_tab = nullptr; int xf_silu = 0;
};

// LayerNorm row partials: 80-column parts, whatever tile width produced them.
inline int gemm_row_parts(int n_out) { return 2 * ((n_out + 159) / 160); }

inline int largest_pow2_divisor(int x, int cap) {
  int p = 1;
  while (p * 2 <= cap && x % (p * 2) == 0) p *= 2;
  return p;
}

// Fused GroupNorm statistics: every epilogue warp (32 accumulator rows) must cover one 32-pixel slab of one sample, and a
// warp's 80/160-column half must hold whole gn_blk-channel blocks.  W = H = 0 for plain (row-major) GEMMs.
inline bool gn_stats_supported(int hw, int W, int H, int blk, int n_out) {
  if (blk <= 0 || blk % 2 || 80 % blk || n_out % blk || hw % 32) return false;
  if (W > 0) {
    const int bw = largest_pow2_divisor(W, 128), bh = largest_pow2_divisor(H, 128 / bw);
    if ((bw * bh) % 32) return false;
  }
  return true;
}

// K-split factor.  Cost model in units of one k-block of MMA time (~0.33 us): a tile costs kb + epilogue, a split adds the
// partial-tile reduction through L2 (measured ~1.6 us per MB of fp32 vector reductions -- profiles/r01_splitk.txt), so
// splitting only pays for long-K, few-tile layers (the bottom of the U).  Every split gets >= 16 k-blocks and the whole
// launch stays within one wave of the persistent grid.  DG_SPLITS=n forces a factor (experiments).
inline int choose_splits(int units, int num_kb, int slots, size_t ws_floats_per_unit, size_t ws_cap) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("DG_SPLITS"); forced = e && e[0] ? atoi(e) : 0; }
  if ((size_t)units * ws_floats_per_unit > ws_cap) return 1;
  if (forced > 0) return (forced * 4 <= num_kb) ? forced : 1;
  static int min_total = -1;
  if (min_total < 0) { const char* e = getenv("DG_SK_MINTOTAL"); min_total = e ? atoi(e) : 32; }
  if (units >= slots || num_kb < min_total) return 1;
  int best = 1;
  double best_cost = 1e30;
  const double mb_per_unit = (double)ws_floats_per_unit * 4.0 / 1e6;
  static double coef = -1.0; static int minkb = -1;
  if (coef < 0) { const char* e = getenv("DG_SK_COEF"); coef = e ? atof(e) : 4.8; }
  if (minkb < 0) { const char* e = getenv("DG_SK_MINKB"); minkb = e ? atoi(e) : 16; }
  for (int s = 1; s <= 16 && s * minkb <= num_kb && units * s <= slots; ++s) {
    const double cost = (double)((num_kb + s - 1) / s) + 9.0 + (s > 1 ? 10.0 + coef * mb_per_unit * units * s : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

// Two epilogue sets (gemm2_kernel<..., kSets = 2>) pay off only when a CTA works through many tiles: measured per shape
// (profiles/r02_ab.md) 32768x960x320 41.0 -> 36.0 us, GEGLU 32768x2560x320 89.7 -> 82.7, 8192x5120x640 57.7 -> 52.4, but
// +1..2 us on every launch with <= 4 tiles per CTA (640-thread CTAs, 112-register epilogue, idle second set).
// DG_GEMM_SETS=2: whenever the kernel allows it; default (and 1): never.
// tensor maps of the extra 1x1 sources of the launch being enqueued (set by launch_gemm just before launch_gemm2_t reads them)
inline thread_local CUtensorMap g_mapX[2];

inline bool gemm_two_sets(int total_units, int slots) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DG_GEMM_SETS"); v = (e && e[0]) ? atoi(e) : 0; }
  (void)total_units; (void)slots;
  return v == 2;        // default: one set -- on the whole forward the selective rule (>= 6 tiles per CTA) measured 9.58-9.68 ms
}                       // against 9.53-9.58 ms with one set everywhere (same box, three alternations)

template <int kCta, int kBN, int kStages, bool kGeglu, bool kXf = false, int kSets = 1>
inline cudaError_t launch_gemm2_t(cudaStream_t stream, int grid_ctas, const CUtensorMap& mA0, const CUtensorMap& mA1,
                                  const CUtensorMap& mW, const CUtensorMap& mO, const CUtensorMap& mR, const Gemm2Params& p) {
  using Cfg = Gemm2Cfg<kCta, kBN, kStages, kSets>;
  return launch_pdl(gemm2_kernel<kCta, kBN, kStages, kGeglu, kXf, kSets>, dim3((unsigned)grid_ctas), dim3(Cfg::kThreads),
                    (size_t)Cfg::kTotal, stream, kCta, mA0, mA1, mW, mO, mR, g_mapX[0], g_mapX[1], p);
}

inline int launch_gemm(cudaStream_t stream, const GemmRes& res, const GemmArgs& a) {
  if (a.c0 % 64 || a.c1 % 64 || a.c0 <= 0 || a.cx0 % 64 || a.cx1 % 64) return fail(DG_E_SHAPE, "gemm: channel counts must be multiples of 64 (%d,%d,%d,%d)", a.c0, a.c1, a.cx0, a.cx1);
  if ((a.cx0 || a.cx1) && (a.taps != 9 || a.in_stride != 1 || a.xf_tab || a.geglu || !a.x0 || (a.cx1 && !a.x1)))
    return fail(DG_E_ARG, "gemm: extra 1x1 sources go with the plain stride-1 3x3 conv");
  if (a.taps != 1 && a.taps != 9 && a.taps != 4) return fail(DG_E_ARG, "gemm: taps must be 1, 9 or 4");
  if (a.taps == 4 && (a.residual || a.rowvec || a.xf_tab || a.geglu)) return fail(DG_E_ARG, "gemm: an upsample phase takes bias / GroupNorm sums only");
  if (a.in_stride != 1 && (a.in_stride != 2 || a.taps != 9 || a.xf_tab || a.c1)) return fail(DG_E_ARG, "gemm: in_stride 2 is the single-source stride-2 3x3 conv");
  if ((reinterpret_cast<uintptr_t>(a.a0) | reinterpret_cast<uintptr_t>(a.w) | reinterpret_cast<uintptr_t>(a.out) |
       reinterpret_cast<uintptr_t>(a.residual) | reinterpret_cast<uintptr_t>(a.a1) | reinterpret_cast<uintptr_t>(a.rowvec)) & 15)
    return fail(DG_E_ARG, "gemm: pointers must be 16-byte aligned");
  if (a.ldo % 8 || (a.residual && a.ld_res % 8) || (a.rowvec && a.ld_rowvec % 8))
    return fail(DG_E_SHAPE, "gemm: output / residual / rowvec row pitch must be a multiple of 8 elements");
  if (a.geglu && (a.residual || a.rowvec || a.row_stats_out || a.gn_stats_out)) return fail(DG_E_ARG, "gemm: geglu epilogue takes bias / LayerNorm fold only");
  if (a.colsum && (!a.ln_stats || !a.bias32 || a.ln_c <= 0)) return fail(DG_E_ARG, "gemm: LayerNorm fold needs ln_stats, bias32 and ln_c");
  if (a.row_stats_out && a.taps != 1) return fail(DG_E_ARG, "gemm: row statistics are produced by plain GEMMs only");
  if (a.xf_tab && (a.geglu || res.cta_mode == 1)) return fail(DG_E_ARG, "gemm: the fused GroupNorm transform is built for CTA-pair, non-GEGLU tiles");
  if (a.gn_stats_out && !gn_stats_supported(a.taps == 1 ? (a.hw > 0 ? a.hw : a.H * a.W) : a.H * a.W, a.taps != 1 ? a.W : 0,
                                            a.taps != 1 ? a.H : 0, a.gn_blk, a.n_out))
    return fail(DG_E_SHAPE, "gemm: fused GroupNorm statistics unsupported for this shape (hw %d, blk %d, n_out %d)", a.H * a.W, a.gn_blk, a.n_out);
  const int kcta = res.cta_mode == 1 ? 1 : 2;
  Gemm2Params p{};
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
  p.hw = a.taps == 1 ? (a.hw > 0 ? a.hw : a.H * a.W) : 0;
  p.n_gemm = a.n_w;
  const int num_kb = a.taps * (a.c0 / 64 + a.c1 / 64) + a.cx0 / 64 + a.cx1 / 64;
  // tile width: 320 (two accumulators, single TMEM stage: best operand reuse) for long-K layers; 160 (double-buffered TMEM:
  // the epilogue overlaps the next main loop, twice the tiles for wave balance) for short-K / epilogue-bound layers
  static int forced_bn = -1;
  if (forced_bn < 0) { const char* e = getenv("DG_GEMM_BN"); forced_bn = e && e[0] ? atoi(e) : 0; }
  static int kb_thresh = -1;
  // round 2: 4 (was 24).  Layers with K = 320..1536 and N > 160 are operand-supply-bound, not epilogue-bound: a 320-wide tile
  // reads its activation rows once instead of once per 160-wide column tile (forward 9.57 -> 9.35-9.46 ms, profiles/r02_ab.md)
  if (kb_thresh < 0) { const char* e = getenv("DG_GEMM_KB_THRESH"); kb_thresh = e ? atoi(e) : 4; }
  int kbn = a.geglu ? kGegluTile : (num_kb > kb_thresh && a.n_w > 160) ? 320 : 160;
  {
    // few-tile layers (the 8x8 / 16x16 levels): 320-wide tiles would leave most SMs idle -- measured on 512x11520x1280:
    // 58 us (320, 4 splits) vs 43 us (160, 4 splits); 2048x11520x1280: 84 vs 62 us (profiles/r01_ncu_summary.md)
    const int m_tiles_ = p.tiles_x * p.tiles_y * p.tiles_b;
    const int units320 = ((m_tiles_ + kcta - 1) / kcta) * ((a.n_w + 319) / 320);
    const int slots_ = kcta == 2 ? res.max_pairs : res.num_sms;
    if (!a.geglu && kbn == 320 && units320 * 2 <= slots_) kbn = 160;
    // K <= 1536: the 320-wide tile's only advantage is reading its activation rows once (~15 % per tile, measured on
    // 32768x960x320 and 32768x320x1280); it loses when the tile count quantises badly against the persistent grid
    // (2048x3840x1280: 96 tiles on 74 CTA pairs, 30.1 vs 26.5 us)
    if (!a.geglu && kbn == 320 && num_kb <= 24) {
      const int units160 = ((m_tiles_ + kcta - 1) / kcta) * ((a.n_w + 159) / 160);
      const int w320 = (units320 + slots_ - 1) / slots_, w160 = (units160 + slots_ - 1) / slots_;
      if (1.7 * w320 > (double)w160) kbn = 160;
    }
  }
  if (forced_bn == 160 || forced_bn == 320) kbn = a.geglu ? kGegluTile : forced_bn;
  if (a.row_stats_out && a.geglu) return fail(DG_E_ARG, "gemm: row statistics are not produced by GEGLU tiles");
  p.tiles_n = (a.n_w + kbn - 1) / kbn;
  p.n_out = a.n_out;
  p.taps = a.taps;
  p.kb0 = a.c0 / 64; p.kb1 = a.c1 / 64; p.kbx0 = a.cx0 / 64; p.kbx1 = a.cx1 / 64;
  p.bias = a.bias; p.bias32 = a.bias32; p.colsum = a.colsum;
  p.ln_stats = a.ln_stats; p.ln_parts = a.ln_parts; p.ln_inv_c = a.ln_c > 0 ? 1.0f / (float)a.ln_c : 0.f; p.ln_eps = a.ln_eps;
  p.rowvec = a.rowvec; p.ld_rowvec = a.ld_rowvec;
  p.residual = a.residual; p.ld_res = a.ld_res;
  p.row_stats_out = a.row_stats_out; p.row_parts = gemm_row_parts(a.n_out);
  p.gn_stats_out = a.gn_stats_out; p.gn_blk = a.gn_blk; p.gn_nblk = a.gn_blk > 0 ? a.n_out / a.gn_blk : 0;
  p.gn_slots = (a.taps == 1 ? (a.hw > 0 ? a.hw : a.H * a.W) : a.H * a.W) / 32;
  p.gn_slot0 = 0; p.tap_x0 = p.tap_y0 = 0; p.out_mul = 1; p.out_ox = p.out_oy = 0; p.in_mul = a.in_stride;
  if (a.taps == 4) {
    p.tap_x0 = a.phase_x - 1; p.tap_y0 = a.phase_y - 1; p.out_mul = 2; p.out_ox = a.phase_x; p.out_oy = a.phase_y;
    if (a.gn_stats_out) { p.gn_slot0 = a.gn_slot0; p.gn_slots = a.gn_slots; }
  }
  p.ws = res.ws; p.tickets = res.tickets;
  p.xf_tab = a.xf_tab; p.xf_c = a.c0 + a.c1; p.xf_silu = a.xf_silu;
  p.xf_one = (a.taps == 1) ? (p.hw > 0 && p.hw % 128 == 0) : (p.bn == 1);
  { int l = 0; while ((1 << l) < p.bw) ++l; p.bw_log2 = l; l = 0; while ((1 << l) < p.bh) ++l; p.bh_log2 = l; }
  if (a.geglu && !a.bias && !a.bias32) return fail(DG_E_ARG, "gemm: geglu epilogue needs a packed bias");

  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int m_units = (m_tiles + kcta - 1) / kcta;
  const int units = m_units * p.tiles_n;
  const int slots = kcta == 2 ? res.max_pairs : res.num_sms;
  p.splits = 1;
  if (res.ws && res.tickets && !a.row_stats_out && m_tiles * p.tiles_n <= kSplitTickets)
    p.splits = choose_splits(units, num_kb, slots, (size_t)kcta * 128 * kbn, kSplitWsFloats);

  p.sk_bulk = (p.splits > 1 && units * p.splits <= slots) ? 1 : 0;   // one tile per CTA: operand stages are idle for the final sum
  { static int nb = -1; if (nb < 0) { const char* e = getenv("DG_SK_BULK"); nb = (e && e[0] == '0') ? 0 : 1; } if (!nb) p.sk_bulk = 0; }
  p.inv_splits = 1.0f / (float)p.splits; p.inv_tiles_n = 1.0f / (float)p.tiles_n;
  p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  CUtensorMap mA0, mA1, mW, mO, mR;
  {
    const uint64_t im = (uint64_t)a.in_stride;    // stride-2 conv: the source is the [B, 2H, 2W, c0] tensor, every second pixel per box
    uint64_t dims[4] = {(uint64_t)a.c0, (uint64_t)W * im, (uint64_t)H * im, (uint64_t)B};
    uint64_t st[3] = {(uint64_t)a.c0 * 2, (uint64_t)W * im * a.c0 * 2, (uint64_t)H * im * W * im * a.c0 * 2};
    uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    DG_TRY(make_map_4d(&mA0, a.a0, dims, st, box, false, false, (uint32_t)im));
    if (a.c1 > 0) {
      uint64_t d1[4] = {(uint64_t)a.c1, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      uint64_t s1[3] = {(uint64_t)a.c1 * 2, (uint64_t)W * a.c1 * 2, (uint64_t)H * W * a.c1 * 2};
      DG_TRY(make_map_4d(&mA1, a.a1, d1, s1, box));
    } else {
      mA1 = mA0;
    }
    g_mapX[0] = mA0; g_mapX[1] = mA0;
    if (a.cx0) {
      uint64_t dx_[4] = {(uint64_t)a.cx0, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      uint64_t sx_[3] = {(uint64_t)a.cx0 * 2, (uint64_t)W * a.cx0 * 2, (uint64_t)H * W * a.cx0 * 2};
      DG_TRY(make_map_4d(&g_mapX[0], a.x0, dx_, sx_, box));
    }
    if (a.cx1) {
      uint64_t dx_[4] = {(uint64_t)a.cx1, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      uint64_t sx_[3] = {(uint64_t)a.cx1 * 2, (uint64_t)W * a.cx1 * 2, (uint64_t)H * W * a.cx1 * 2};
      DG_TRY(make_map_4d(&g_mapX[1], a.x1, dx_, sx_, box));
    }
    const uint64_t ktot = (uint64_t)a.taps * (a.c0 + a.c1) + a.cx0 + a.cx1;
    DG_TRY(make_map_2d(&mW, a.w, ktot, (uint64_t)a.n_w, ktot * 2, 64, (kbn == 256 ? 128 : 160) / kcta));   // one accumulator's rows per CTA
    // output tiles: {n_out, W, H, B} boxes of 32 columns x 128 pixels, 64-byte swizzle in smem
    const uint64_t om = a.taps == 4 ? 2 : 1;      // an upsample phase scatters into the [B, 2H, 2W, n_out] output
    uint64_t od[4] = {(uint64_t)a.n_out, (uint64_t)W * om, (uint64_t)H * om, (uint64_t)B};
    uint64_t os[3] = {(uint64_t)a.ldo * 2, (uint64_t)W * om * a.ldo * 2, (uint64_t)H * om * W * om * a.ldo * 2};
    uint32_t obox[4] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    DG_TRY(make_map_4d(&mO, a.out, od, os, obox, false, true, (uint32_t)om));
    if (a.residual) {     // same boxes over the residual tensor: the TMA load lands where the TMA store will read
      uint64_t rs[3] = {(uint64_t)a.ld_res * 2, (uint64_t)W * a.ld_res * 2, (uint64_t)H * W * a.ld_res * 2};
      DG_TRY(make_map_4d(&mR, a.residual, od, rs, obox, false, true));
    } else {
      mR = mO;
    }
  }
  const int total_units = units * p.splits;
  const int grid_units = total_units < slots ? total_units : slots;
  {
    // DG_GEMM_BLOCKED=0: every GEMM keeps the strided tile assignment; DG_GEMM_VECPRE=0: per-column vectors loaded at the tile's start
    static int blocked = -1, vecpre = -1;
    if (blocked < 0) { const char* e = getenv("DG_GEMM_BLOCKED"); blocked = e ? atoi(e) : 1; }   // 2: every multi-column-tile GEMM, not only the LayerNorm-folded ones
    if (vecpre < 0) { const char* e = getenv("DG_GEMM_VECPRE"); vecpre = (e && e[0] == '0') ? 0 : 1; }
    p.vec_pre = vecpre;
    p.blk_np = 0; p.blk_q = 0; p.blk_rem = 0; p.inv_blk_np = 0.f;
    if (blocked && (a.ln_stats || blocked >= 2) && p.splits == 1 && p.tiles_n > 1 && total_units > grid_units) {
      p.blk_np = grid_units; p.blk_q = total_units / grid_units; p.blk_rem = total_units % grid_units;
      p.inv_blk_np = 1.0f / (float)grid_units;
    }
  }
  if (trace_on())
    fprintf(stderr, "DG_TRACE gemm M=%d N=%d K=%d taps=%d c0=%d c1=%d bn=%d units=%d splits=%d grid=%dx%d geglu=%d ln=%d res=%d gn=%d rs=%d\n",
            a.B * a.H * a.W, a.n_w, a.taps * (a.c0 + a.c1) + a.cx0 + a.cx1, a.taps, a.c0, a.c1, kbn, units, p.splits, grid_units, kcta, a.geglu,
            a.colsum != nullptr, a.residual != nullptr, a.gn_stats_out != nullptr, a.row_stats_out != nullptr);
  const double rows_ = (double)a.B * a.H * a.W, ktot_ = (double)a.taps * (a.c0 + a.c1) + a.cx0 + a.cx1;
  // algorithmic work: a GEGLU projection has 2 * n_out weight rows (the packing's zero rows are not work)
  const double n_alg_ = a.geglu ? 2.0 * a.n_out : (double)a.n_w;
  ProfScope prof_(FAM_GEMM, stream, 2.0 * rows_ * ktot_ * n_alg_,
                  2.0 * (rows_ * (a.c0 + a.c1 + a.cx0 + a.cx1) + ktot_ * n_alg_ + rows_ * a.n_out));
  static int dbg_on = -1;
  static long long* dbg_dev = nullptr;
  if (dbg_on < 0) {
    const char* ev = getenv("DG_GEMM_DBG"); dbg_on = ev && ev[0] == '1';
#ifndef DG_GEMM_STAMPS
    if (dbg_on) { fprintf(stderr, "DG_GEMM_DBG needs a library built with DG_NVCC_EXTRA=-DDG_GEMM_STAMPS\n"); dbg_on = 0; }
#endif
    if (dbg_on) { cudaMalloc(&dbg_dev, 40 * 8); }
  }
  if (dbg_on) { cudaMemsetAsync(dbg_dev, 0, 40 * 8, stream); p.dbg = dbg_dev; }
  cudaError_t e;
  if (kcta == 2) {
    if (a.geglu && kGegluSets == 2 && p.splits == 1 && gemm_two_sets(total_units, slots))
      e = launch_gemm2_t<2, kGegluTile, kStagesGeglu2, true, false, kGegluSets>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
    else if (a.geglu) e = launch_gemm2_t<2, kGegluTile, kStagesGeglu2, true>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
    else if (a.xf_tab && kbn == 320) e = launch_gemm2_t<2, 320, kStages_2_320, false, true>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
    else if (a.xf_tab) e = launch_gemm2_t<2, 160, kStages_2_160, false, true>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
    else if (kbn == 320) e = launch_gemm2_t<2, 320, kStages_2_320, false>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
    // 160-wide tiles without split-K, many tiles per CTA: two epilogue sets (16 epilogue warps, two tiles' epilogues in flight)
    else if (p.splits == 1 && gemm_two_sets(total_units, slots)) e = launch_gemm2_t<2, 160, kStages_2_160, false, false, 2>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
    else e = launch_gemm2_t<2, 160, kStages_2_160, false>(stream, grid_units * 2, mA0, mA1, mW, mO, mR, p);
  } else {
    if (a.geglu) e = launch_gemm2_t<1, kGegluTile, kStagesGeglu1, true>(stream, grid_units, mA0, mA1, mW, mO, mR, p);
    else if (kbn == 320) e = launch_gemm2_t<1, 320, kStages_1_320, false>(stream, grid_units, mA0, mA1, mW, mO, mR, p);
    else e = launch_gemm2_t<1, 160, kStages_1_160, false>(stream, grid_units, mA0, mA1, mW, mO, mR, p);
  }
  ++g_launch_counter;
  if (dbg_on && e == cudaSuccess) {
    long long h[40];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "DG_GEMM_DBG bn=%d units=%d splits=%d: prologue %lld | operands +%lld | acc +%lld | chunks", kbn, units, p.splits,
            h[1] - h[0], h[2] - h[0], h[3] - h[0]);
    for (int i = 4; i < 9; ++i) fprintf(stderr, " +%lld", h[i] - h[0]);
    fprintf(stderr, " | chunk1: ld %lld math %lld bufwait %lld st %lld fence %lld", h[17] - h[16], h[18] - h[17], h[19] - h[18], h[20] - h[19], h[21] - h[20]);
    fprintf(stderr, " | units (acc, handed):");
    for (int i = 0; i < 4; ++i) if (h[22 + 2 * i]) fprintf(stderr, " (+%lld, +%lld)", h[22 + 2 * i] - h[0], h[23 + 2 * i] - h[0]);
    fprintf(stderr, " | tile 1: loop top +%lld, setup done +%lld", h[30] - h[0], h[31] - h[0]);
    fprintf(stderr, " | producer (tile k loads go out):");
    for (int i = 0; i < 4; ++i) if (h[32 + i]) fprintf(stderr, " +%lld", h[32 + i] - h[0]);
    fprintf(stderr, " | MMA (tile k last issue):");
    for (int i = 0; i < 4; ++i) if (h[36 + i]) fprintf(stderr, " +%lld", h[36 + i] - h[0]);
    fprintf(stderr, " | stop +%lld | stores done +%lld | at teardown +%lld | passed +%lld | tmem freed +%lld\n", h[10] - h[0],
            h[11] - h[0], h[12] - h[0], h[13] - h[0], h[14] - h[0]);
  }
  if (e != cudaSuccess) return fail(DG_E_CUDA, "gemm2 launch failed: %s", cudaGetErrorString(e));
  return DG_OK;
}

// ------------------------------------------------------------------ attention launcher
// Wave-balanced attention grid.  Every attention CTA owns an SM (shared memory + all 512 TMEM columns), so a grid of n_bh * T / big
// equal CTAs (T = 128-row query tiles per (batch, head) pair, big = query tiles per CTA) runs in ceil(ctas / SMs) waves: SD-1.5's
// 64 x 64 level at UNet batch 8 is 512 CTAs = 3.46 waves, i.e. 16 tile times where 2048 tiles / 148 SMs = 13.84 would do.  The plan
// cuts each pair into a CTAs of `big` tiles and b of `big - 1` (big * a + (big - 1) * b = T; two groups of pairs with their own
// (a, b)), orders the grid longest-first and simulates the block scheduler (next CTA -> first free SM); it is used only when the
// simulated makespan beats the uniform grid by >= 3 % (a short CTA is charged 4 % extra per tile).  DG_ATTN_PART=0: off.
inline int attn_part_min_keys() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DG_ATTN_PART_MIN_KEYS"); v = e ? atoi(e) : 512; }
  return v;
}
struct AttnPlan { int on = 0, n_g1 = 0, a1 = 0, b1 = 0, a2 = 0, b2 = 0, ctas = 0; double makespan = 0, uniform = 0; };
inline double attn_plan_makespan(int sms, int n_big, int n_small, double c_big, double c_small) {
  std::vector<double> heap(sms, 0.0);                       // min-heap of SM free times
  auto cmp = [](double x, double y) { return x > y; };
  double end = 0;
  for (int i = 0; i < n_big + n_small; ++i) {
    std::pop_heap(heap.begin(), heap.end(), cmp);
    double t = heap.back() + (i < n_big ? c_big : c_small);
    heap.back() = t;
    end = t > end ? t : end;
    std::push_heap(heap.begin(), heap.end(), cmp);
  }
  return end;
}
inline AttnPlan plan_attn_grid(int n_bh, int T, int big, int sms) {
  static std::map<std::tuple<int, int, int, int>, AttnPlan> cache;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("DG_ATTN_PART"); enabled = (e && e[0] == '0') ? 0 : 1; }
  AttnPlan best;
  if (!enabled || big < 2 || T < big || sms <= 0) return best;
  const auto key = std::make_tuple(n_bh, T, big, sms);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const int small = big - 1;
  const double c_big = big, c_small = small * 1.04;
  const long uniform_ctas = (long)n_bh * ((T + big - 1) / big);
  best.uniform = (double)((uniform_ctas + sms - 1) / sms) * c_big;
  best.makespan = best.uniform;
  std::vector<std::pair<int, int>> opts;                    // (a, b): the (up to 6) splits of a pair with the most full-size CTAs
  for (int a = T / big; a >= 0 && opts.size() < 6; --a)
    if ((T - big * a) % small == 0) opts.push_back({a, (T - big * a) / small});
  for (auto& o1 : opts)
    for (auto& o2 : opts)
      for (int g1 = (o1 == o2 ? n_bh : 1); g1 <= n_bh; ++g1) {
        if (o1 == o2 && g1 != n_bh) continue;
        if (o1 != o2 && g1 == n_bh) continue;               // (covered by o1 == o2)
        const int g2 = n_bh - g1;
        const int nb = g1 * o1.first + g2 * o2.first, ns = g1 * o1.second + g2 * o2.second;
        const double m = attn_plan_makespan(sms, nb, ns, c_big, c_small);
        if (m < best.makespan - 1e-9 || (best.on && m < best.makespan + 1e-9 && nb + ns < best.ctas)) {
          best.on = 1; best.makespan = m; best.n_g1 = g1; best.a1 = o1.first; best.b1 = o1.second; best.a2 = o2.first; best.b2 = o2.second;
          best.ctas = nb + ns;
        }
      }
  if (best.on && best.makespan > 0.97 * best.uniform) best.on = 0;
  cache[key] = best;
  return best;
}

template <int kD, int kKV, int kStages, int kSBuf, int kQ = 2, int kSplit = 1>
inline int launch_attn_t(cudaStream_t stream, const __half* q, int ldq, const __half* k, int ldk, const __half* v,
                         int ldv, __half* out, int B, int heads, int Sq, int Sk, int causal = 0) {
  using C = AttnCfg<kD, kKV, kStages, kSBuf, kQ, kSplit>;
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
  p.Sq = Sq; p.Sk = Sk; p.heads = heads; p.ldo = heads * kD; p.out = out; p.causal = causal;
  if (causal && kSplit != 1) return fail(DG_E_UNSUPPORTED, "attention: causal mask needs a one-stream-per-query-tile variant");
  p.scale_log2 = (1.0f / sqrtf((float)kD)) * 1.4426950408889634f;
  static int poly = -1;
  if (poly < 0) { const char* e = getenv("DG_ATTN_POLY"); poly = e ? atoi(e) : 3; }
  auto kern = poly == 0 ? attn_tc_kernel<kD, kKV, kStages, 0, kSBuf, kQ, kSplit>
              : poly == 2 ? attn_tc_kernel<kD, kKV, kStages, 2, kSBuf, kQ, kSplit> : attn_tc_kernel<kD, kKV, kStages, 3, kSBuf, kQ, kSplit>;
  dim3 grid((Sq + C::kQTiles * 128 - 1) / (C::kQTiles * 128), heads, B);
  if (kSplit == 1 && !causal && Sq % 128 == 0 && Sk >= attn_part_min_keys()) {   // (short key sequences: a CTA's time is its fixed cost, not its tiles)
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const AttnPlan pl = plan_attn_grid(B * heads, Sq / 128, C::kQTiles, sms);
    if (pl.on) {
      p.part_on = 1; p.n_bh = B * heads; p.n_g1 = pl.n_g1; p.a1 = pl.a1; p.b1 = pl.b1; p.a2 = pl.a2; p.b2 = pl.b2;
      grid = dim3(pl.ctas, 1, 1);
    }
  }
  if (trace_on()) fprintf(stderr, "DG_TRACE attn B=%d heads=%d Sq=%d Sk=%d d=%d grid=%u (%s)\n", B, heads, Sq, Sk, kD, grid.x * grid.y * grid.z,
                          p.part_on ? "wave-balanced" : "uniform");
  ProfScope prof_(FAM_ATTN, stream, 4.0 * B * heads * (double)Sq * Sk * kD,
                  2.0 * B * heads * kD * (2.0 * Sq + 2.0 * Sk));
  {
    ++g_launch_counter;
    cudaError_t e = launch_pdl(kern, grid, dim3(C::kThreads), (size_t)C::kSmem, stream, 1, mQ, mK, mV, p);
    if (e != cudaSuccess) return fail(DG_E_CUDA, "attention launch failed: %s", cudaGetErrorString(e));
  }
  return DG_OK;
}

inline bool attn_x96() {
  // measured SLOWER (S = 4096 x 77, d = 40: 42.2 vs 36.6 us; forward 9.67-9.73 vs 9.61-9.64 ms): 8 softmax warps on the masked
  // per-element path lose more than the second, mostly empty key tile costs.  DG_ATTN_X96=1 enables it.
  static int v = -1;
  if (v < 0) { const char* e = getenv("DG_ATTN_X96"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

inline int launch_attention(cudaStream_t stream, const __half* q, int ldq, const __half* k, int ldk, const __half* v,
                            int ldv, __half* out, int B, int heads, int Sq, int Sk, int d, int causal = 0) {
  if (causal && d != 64) return fail(DG_E_UNSUPPORTED, "attention: causal mask is built for head dim 64 only");
  if ((ldq | ldk | ldv) % 8) return fail(DG_E_SHAPE, "attention: row strides must be multiples of 8 elements");
  switch (d) {
    case 32: return launch_attn_t<32, 128, 4, 1>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    case 40: {
      // cross-attention (77 text tokens): ONE 96-key tile instead of two 64-key tiles (64 + 13 valid keys); DG_ATTN_X96=0: off
      if (attn_x96() && Sk <= 96 && !causal) return launch_attn_t<40, 96, 2, 1, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      static int var = -1;     // DG_ATTN_VAR: 0 = 2 query tiles x 128 keys, 1 = 2 x 64 keys double-buffered scores, 2 = 4 query tiles x 64 keys
      if (var < 0) { const char* e = getenv("DG_ATTN_VAR"); var = e ? atoi(e) : 2; }   // 3 = 2 query tiles x 2 key-tile streams (measured equal to 2 at batch 8; better wave fit at small batch)
      if (var == 0) return launch_attn_t<40, 128, 4, 1>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      if (var == 1) return launch_attn_t<40, 64, 6, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      if (var == 2) return launch_attn_t<40, 64, 6, 1, 4>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      if (var == 4) return launch_attn_t<40, 32, 12, 2, 4>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);   // 4 query tiles x 32 keys, double-buffered scores (fills TMEM exactly)
      return launch_attn_t<40, 64, 6, 1, 4, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    }
    case 64: {
      // DG_ATTN64_VAR: 0 = 2 query tiles x 128 keys (8 softmax warps), 1 = 4 query tiles x 64 keys (16 softmax warps; the d = 40
      // kernel's best shape).  Causal (CLIP text, 77 tokens) and short sequences stay on variant 0.
      static int var64 = -1;
      if (var64 < 0) { const char* e = getenv("DG_ATTN64_VAR"); var64 = e ? atoi(e) : 0; }
      if (attn_x96() && Sk <= 96 && !causal && Sq >= 256) return launch_attn_t<64, 96, 2, 1, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      if (var64 == 2 && !causal && Sq >= 512)     // 2 query tiles x 64 keys, double-buffered scores
        return launch_attn_t<64, 64, 6, 2>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      if (var64 == 1 && !causal && Sq >= 512)
        return launch_attn_t<64, 64, 6, 1, 4>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
      return launch_attn_t<64, 128, 3, 1>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk, causal);
    }
    case 80: return launch_attn_t<80, 128, 2, 1>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    case 160: return launch_attn_t<160, 64, 2, 1>(stream, q, ldq, k, ldk, v, ldv, out, B, heads, Sq, Sk);
    default: return fail(DG_E_UNSUPPORTED, "attention: head dim %d not built (32/40/64/80/160)", d);
  }
}

// Opt every tcgen05 kernel into its dynamic shared-memory size once per device (never during graph capture).
template <int kD, int kKV, int kStages, int kSBuf = 1, int kQ = 2, int kSplit = 1>
inline int init_attn_attr() {
  DG_CUDA(cudaFuncSetAttribute(attn_tc_kernel<kD, kKV, kStages, 0, kSBuf, kQ, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<kD, kKV, kStages, kSBuf, kQ, kSplit>::kSmem));
  DG_CUDA(cudaFuncSetAttribute(attn_tc_kernel<kD, kKV, kStages, 3, kSBuf, kQ, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<kD, kKV, kStages, kSBuf, kQ, kSplit>::kSmem));
  DG_CUDA(cudaFuncSetAttribute(attn_tc_kernel<kD, kKV, kStages, 2, kSBuf, kQ, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<kD, kKV, kStages, kSBuf, kQ, kSplit>::kSmem));
  return DG_OK;
}
template <int kCta, int kBN, int kStages, bool kGeglu, bool kXf = false, int kSets = 1>
inline int init_gemm_attr() {
  DG_CUDA(cudaFuncSetAttribute(gemm2_kernel<kCta, kBN, kStages, kGeglu, kXf, kSets>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Gemm2Cfg<kCta, kBN, kStages, kSets>::kTotal));
  return DG_OK;
}
// Co-resident CTA pairs of the pair kernels (persistent grid size; every variant is one CTA per SM).
inline int query_max_pairs(int* out) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * 148); cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = Gemm2Cfg<2, 160, kStages_2_160>::kTotal;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  DG_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<2, 160, kStages_2_160, false>, &cfg));
  if (n <= 0) return fail(DG_E_CUDA, "no co-resident CTA pair fits (cudaOccupancyMaxActiveClusters = %d)", n);
  *out = n;
  return DG_OK;
}
inline int init_kernel_attributes() {
  DG_TRY((init_gemm_attr<1, 320, kStages_1_320, false>()));
  DG_TRY((init_gemm_attr<1, 160, kStages_1_160, false>()));
  DG_TRY((init_gemm_attr<1, kGegluTile, kStagesGeglu1, true>()));
  DG_TRY((init_gemm_attr<2, kGegluTile, kStagesGeglu2, true>()));
  if constexpr (kGegluSets == 2) DG_TRY((init_gemm_attr<2, kGegluTile, kStagesGeglu2, true, false, kGegluSets>()));
  DG_TRY((init_gemm_attr<2, 320, kStages_2_320, false>()));
  DG_TRY((init_gemm_attr<2, 160, kStages_2_160, false>()));
  DG_TRY((init_gemm_attr<2, 320, kStages_2_320, false, true>()));
  DG_TRY((init_gemm_attr<2, 160, kStages_2_160, false, true>()));
  DG_TRY((init_gemm_attr<2, 160, kStages_2_160, false, false, 2>()));
  DG_TRY((init_attn_attr<32, 128, 4>()));
  DG_TRY((init_attn_attr<40, 128, 4>()));
  DG_TRY((init_attn_attr<40, 64, 6, 2>()));
  DG_TRY((init_attn_attr<40, 64, 6, 1, 4>()));
  DG_TRY((init_attn_attr<40, 32, 12, 2, 4>()));
  DG_TRY((init_attn_attr<40, 64, 6, 1, 4, 2>()));
  DG_TRY((init_attn_attr<64, 128, 3>()));
  DG_TRY((init_attn_attr<64, 64, 6, 1, 4>()));
  DG_TRY((init_attn_attr<64, 64, 6, 2>()));
  DG_TRY((init_attn_attr<40, 96, 2, 1, 2>()));
  DG_TRY((init_attn_attr<64, 96, 2, 1, 2>()));
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

inline void gn_launch_geometry(int C, int B, int HW, int num_sms, int* ppb_out, int* pstride_out) {
  const int nvec = C / 8;
  const int pstride = nvec <= 256 ? 256 / nvec : 1;
  // strip length: enough blocks to fill the chip (>= 4 per SM when the tensor allows), >= 8 pixel iterations per thread
  static int it0 = -1, waves = -1;
  if (it0 < 0) { const char* e = getenv("DG_GN_ITERS"); it0 = e && e[0] ? atoi(e) : 8; const char* w = getenv("DG_GN_WAVES"); waves = w && w[0] ? atoi(w) : 8; }
  int ppb = it0 * pstride;
  while (ppb * 2 <= HW && (size_t)B * ((HW + ppb - 1) / ppb) > (size_t)waves * num_sms) ppb *= 2;
  *ppb_out = ppb; *pstride_out = pstride;
}

inline int launch_groupnorm(cudaStream_t s, int num_sms, const __half* x0, int C0, const __half* x1, int C1,
                            const __half* gamma, const __half* beta, __half* out, float* stats, int B, int HW,
                            int groups, float eps, int silu) {
  const int C = C0 + C1;
  if (C % groups || C0 % 8 || C1 % 8) return fail(DG_E_SHAPE, "groupnorm: C=%d+%d groups=%d", C0, C1, groups);
  ProfScope prof_(FAM_NORM, s, 0.0, 2.0 * 3.0 * B * HW * (double)C);
  DG_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * B, s));
  int ppb, pstride;
  gn_launch_geometry(C, B, HW, num_sms, &ppb, &pstride);
  dim3 grid((HW + ppb - 1) / ppb, B);
  const size_t smem = sizeof(float) * 2 * (size_t)pstride * C;
  gn_stats_kernel<<<grid, 256, smem, s>>>(x0, C0, x1, C1, HW, groups, ppb, stats);
  DG_LAUNCH_CHECK();
  gn_apply_kernel<<<grid, 256, 0, s>>>(x0, C0, x1, C1, HW, groups, eps, ppb, stats, gamma, beta, silu, out);
  DG_LAUNCH_CHECK();
  return DG_OK;
}

// GroupNorm apply with statistics already accumulated by the producing GEMM epilogues (block sums per source).
inline int launch_groupnorm_fused(cudaStream_t s, int num_sms, const __half* x0, int C0, const float* st0, const __half* x1,
                                  int C1, const float* st1, int blk, const __half* gamma, const __half* beta, __half* out,
                                  int B, int HW, int groups, float eps, int silu, float* group_totals = nullptr) {
  if (HW % 32) return fail(DG_E_SHAPE, "groupnorm(fused): HW=%d must be a multiple of 32", HW);
  const int C = C0 + C1;
  if (C % groups || C0 % 8 || C1 % 8 || groups > 64) return fail(DG_E_SHAPE, "groupnorm: C=%d+%d groups=%d", C0, C1, groups);
  if (blk <= 0 || (C / groups) % blk || C0 % blk || C1 % blk || C / blk > 256)
    return fail(DG_E_SHAPE, "groupnorm: block %d does not tile C=%d+%d (at most 256 blocks)", blk, C0, C1);
  ProfScope prof_(FAM_NORM, s, 0.0, 2.0 * 2.0 * B * HW * (double)C);
  int ppb, pstride;
  gn_launch_geometry(C, B, HW, num_sms, &ppb, &pstride);
  dim3 grid((HW + ppb - 1) / ppb, B);
  // DG_GN_TWO_STEP=1: fold the block sums once (one CTA per sample) into group totals, then the lean apply kernel -- instead of
  // every apply CTA folding its sample's slabs x blocks itself (as many bytes as the activations it then touches)
  static int two_step = -1;
  if (two_step < 0) { const char* e = getenv("DG_GN_TWO_STEP"); two_step = (e && e[0] == '1') ? 1 : 0; }
  if (two_step && group_totals) {
    g_launch_counter += 2;
    cudaError_t e = launch_pdl(gn_fold_groups_kernel, dim3((unsigned)B), dim3(256), (size_t)0, s, 1, C0, C1, groups, st0, st1, blk, HW / 32,
                               group_totals);
    if (e == cudaSuccess)
      e = launch_pdl(gn_apply_kernel, grid, dim3(256), (size_t)0, s, 1, x0, C0, x1, C1, HW, groups, eps, ppb, (const float*)group_totals, gamma,
                     beta, silu, out);
    if (e != cudaSuccess) return fail(DG_E_CUDA, "groupnorm launch failed: %s", cudaGetErrorString(e));
    return DG_OK;
  }
  {
    ++g_launch_counter;
    cudaError_t e = launch_pdl(gn_apply_blk_kernel, grid, dim3(256), (size_t)0, s, 1, x0, C0, x1, C1, HW, groups, eps, ppb, st0, st1, blk,
                               HW / 32, gamma, beta, silu, out);
    if (e != cudaSuccess) return fail(DG_E_CUDA, "groupnorm launch failed: %s", cudaGetErrorString(e));
  }
  return DG_OK;
}

// Statistics fold for a GroupNorm applied inside the consuming GEMM (GemmArgs::xf_tab): tab = [B][3][C0 + C1] fp16.
inline int launch_gn_fold(cudaStream_t s, int C0, const float* st0, int C1, const float* st1, int blk, const __half* gamma,
                          const __half* beta, __half* tab, int B, int HW, int groups, float eps, int silu) {
  if (HW % 32) return fail(DG_E_SHAPE, "groupnorm(fold): HW=%d must be a multiple of 32", HW);
  const int C = C0 + C1;
  if (C % groups || C0 % 64 || C1 % 64 || groups > 64) return fail(DG_E_SHAPE, "groupnorm(fold): C=%d+%d groups=%d", C0, C1, groups);
  if (blk <= 0 || (C / groups) % blk || C0 % blk || C1 % blk || C / blk > 256)
    return fail(DG_E_SHAPE, "groupnorm(fold): block %d does not tile C=%d+%d (at most 256 blocks)", blk, C0, C1);
  ProfScope prof_(FAM_NORM, s, 0.0, 8.0 * B * (HW / 32) * (double)(C / blk) + 6.0 * B * C);
  ++g_launch_counter;
  cudaError_t e = launch_pdl(gn_fold_kernel, dim3((unsigned)B), dim3(256), (size_t)0, s, 1, C0, C1, HW, groups, eps, st0, st1, blk, HW / 32,
                             gamma, beta, silu, tab);
  if (e != cudaSuccess) return fail(DG_E_CUDA, "groupnorm fold launch failed: %s", cudaGetErrorString(e));
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
