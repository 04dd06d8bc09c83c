// tcgen05 implicit-GEMM kernel: one kernel serves every Linear, 1x1 conv and 3x3 (stride-1, pad-1) conv of the UNet.
//
//   out[pixel, n] = epilogue( sum_{tap, c} A[pixel + tap_offset, c] * Wt[n, tap*C + c] )
//
// A is an NHWC fp16 activation tensor viewed through a 4-D TMA tensor map {C, W, H, B}; a 128-row M tile is a
// {bw x bh x bn} box of pixels, and each 3x3 tap is the same box shifted by (dx, dy) with TMA out-of-bounds zero fill
// supplying the padding.  A plain [M, K] GEMM is the degenerate case {K, M, 1, 1} with one tap.  Up to two A sources
// split the channel range (the `cat([x, skip])` of the up blocks is never materialised for the 1x1 shortcut).
// Wt is the K-major packed weight [N, taps*C] behind a 2-D tensor map.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4..11 = epilogue.  Persistent over tiles, kStages-deep smem ring, two TMEM accumulator stages so the epilogue
// of tile i overlaps the main loop of tile i+1.
//
// Epilogue: the residual tile is TMA-loaded into the (64-byte-swizzled) staging tile while the main loop runs; each
// epilogue thread owns one accumulator row x half the columns: tcgen05.ld -> +bias (fp32, from smem) (+time-embedding
// vector) (+residual, read in place from the staging tile) -> fp16 -> staging tile; one thread then issues TMA stores
// (full-line, clipped at the tensor edge by the tensor map).  No per-element global load/store instructions.
//
// Replaces (reference side): torch.nn.Conv2d / Linear inside diffusers ResnetBlock2D, Attention, FeedForward,
// Transformer2DModel, reached from DiverGen/generation/txt2img_diffusers_stages_from_txt.py:255-259.
#pragma once
#include "common.cuh"

namespace dg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // one 128-byte swizzle atom of fp16

struct GemmParams {
  // problem
  int n_out;        // valid output columns (after GEGLU halving if enabled)
  int n_gemm;       // GEMM N (rows of Wt that are meaningful)
  int taps;         // 1 or 9
  int kb0, kb1;     // 64-channel k-blocks taken from source 0 / source 1 per tap
  // M tiling (pixels)
  int W, H, B;      // logical A dims (plain GEMM: W = M, H = B = 1)
  int bw, bh, bn;   // box; bw*bh*bn == 128
  int tiles_x, tiles_y, tiles_b, tiles_n;
  // epilogue
  const __half* bias;      // [n_gemm] or nullptr
  const __half* rowvec;    // [B, ld_rowvec] per-sample additive vector (time embedding) or nullptr
  int ld_rowvec;
  int has_residual;        // residual tile comes through mapR
};

template <int kBlockN, int kStages, bool kGeglu>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBlockN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutCols = kGeglu ? kBlockN / 2 : kBlockN;
  static constexpr int kSubTiles = kOutCols / 32;            // staging sub-tiles of [128 rows][32 cols], 64B swizzle
  static constexpr int kSubBytes = kBlockM * 64;
  static constexpr int kOutBytes = kSubTiles * kSubBytes;
  static constexpr int kBiasBytes = kBlockN * 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kOutBytes + kBiasBytes + kBarBytes + 1024 /*align slack*/;
  static_assert(kOutCols % 32 == 0, "staging sub-tiles are 32 columns wide");
};

// erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below fp16 resolution): 2 MUFU + ~10 FMA instead of erff().
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.0f + copysignf(e, x));
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct TileCoord { int nt, x0, y0, b0; };
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int tile) {
  TileCoord t;
  t.nt = tile % p.tiles_n;
  int mt = tile / p.tiles_n;
  t.x0 = (mt % p.tiles_x) * p.bw; mt /= p.tiles_x;
  t.y0 = (mt % p.tiles_y) * p.bh;
  t.b0 = (mt / p.tiles_y) * p.bn;
  return t;
}

template <int kBlockN, int kStages, bool kGeglu>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapO,
               const __grid_constant__ CUtensorMap mapR, const GemmParams p) {
  using S = GemmSmem<kBlockN, kStages, kGeglu>;
  static_assert(kBlockN % 32 == 0 && kBlockN <= 256, "UMMA N / epilogue split");
  static_assert(S::kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment");
  constexpr uint32_t kAccStride = 256;  // TMEM columns between the two accumulator stages
  constexpr uint32_t kTmemCols = 512;
  constexpr int kEpiThreads = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOut = smem + kStages * S::kStageBytes;                         // 1024-aligned (stage bytes are)
  float* sBias = reinterpret_cast<float*>(sOut + S::kOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + S::kOutBytes + S::kBiasBytes);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* res_full = acc_empty + 2;       // [1] residual tile landed in the staging tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapA1); tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapO); tma_prefetch_desc(&mapR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiThreads); }
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = m_tiles * p.tiles_n;
  const int kb_per_tap = p.kb0 + p.kb1;
  const int num_kb = p.taps * kb_per_tap;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(p, tile);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dx = (p.taps == 9) ? (tap % 3 - 1) : 0;
          const int dy = (p.taps == 9) ? (tap / 3 - 1) : 0;
          for (int kb = 0; kb < kb_per_tap; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_arrive_expect_tx(&full[stage], S::kStageBytes);
            if (kb < p.kb0) tma_load_4d(sa, &mapA0, &full[stage], kb * kBlockK, tc.x0 + dx, tc.y0 + dy, tc.b0);
            else            tma_load_4d(sa, &mapA1, &full[stage], (kb - p.kb0) * kBlockK, tc.x0 + dx, tc.y0 + dy, tc.b0);
            tma_load_2d(sb, &mapW, &full[stage], (tap * kb_per_tap + kb) * kBlockK, tc.nt * kBlockN);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_f16(kBlockN);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - 4;
    const int q = ew & 3;               // TMEM lane quadrant (== warp index % 4)
    const int hf = ew >> 2;             // which half of the output columns this warp owns
    const int r = q * 32 + lane;        // accumulator row within the tile
    const int et = threadIdx.x - 128;   // 0..255
    const int box_xy = p.bw * p.bh;
    const uint32_t row_sw = (uint32_t)((r >> 1) & 3);   // 64-byte swizzle phase of this row
    uint8_t* my_row = sOut + r * 64;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t tile_iter = 0;

    auto load_residual = [&](int tile) {   // one thread: residual tile -> staging tile (async, signals res_full)
      const TileCoord tc = tile_coord(p, tile);
      mbar_arrive_expect_tx(res_full, S::kOutBytes);
#pragma unroll
      for (int s = 0; s < S::kSubTiles; ++s)
        tma_load_4d(sOut + s * S::kSubBytes, &mapR, res_full, tc.nt * S::kOutCols + s * 32, tc.x0, tc.y0, tc.b0);
    };
    if (et == 0 && p.has_residual && (int)blockIdx.x < total_tiles) load_residual(blockIdx.x);

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
      const TileCoord tc = tile_coord(p, tile);
      const int nt = tc.nt;
      const int bb = min(tc.b0 + r / box_xy, p.B - 1);   // sample of this row (clamped; OOB rows are clipped by TMA)

      // stage this tile's bias as fp32 (its global latency hides behind the accumulator wait)
      float bv = 0.f;
      if (et < kBlockN && p.bias && nt * kBlockN + et < p.n_gemm) bv = __half2float(__ldg(p.bias + nt * kBlockN + et));
      if (et < kBlockN) sBias[et] = bv;   // previous tile's readers are all past the closing bar.sync of that tile
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // sBias visible; thread 0 has seen the previous store drained
      if (p.has_residual) mbar_wait(res_full, tile_iter & 1);
      const uint32_t t_row = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);

      if constexpr (!kGeglu) {
        constexpr int kColsPerThread = kBlockN / 2;
        constexpr int kChunks = kColsPerThread / 16;
        static_assert(kColsPerThread % 16 == 0, "column split");
        uint32_t v[kChunks][16];
#pragma unroll
        for (int j = 0; j < kChunks; ++j) tmem_ld16(t_row + hf * kColsPerThread + j * 16, v[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kChunks; ++j) {
          const int c = hf * kColsPerThread + j * 16;  // column within the tile
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c + i);
            f[i] = __uint_as_float(v[j][i]) + b4.x;         f[i + 1] = __uint_as_float(v[j][i + 1]) + b4.y;
            f[i + 2] = __uint_as_float(v[j][i + 2]) + b4.z; f[i + 3] = __uint_as_float(v[j][i + 3]) + b4.w;
          }
          if (p.rowvec) {
            const int col = nt * kBlockN + c;
            if (col + 16 <= p.n_out) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.rowvec + (size_t)bb * p.ld_rowvec + col);
              const uint4 b0 = __ldg(rp), b1 = __ldg(rp + 1);
              float2 t;
              t = unpack_half2(b0.x); f[0] += t.x; f[1] += t.y;   t = unpack_half2(b0.y); f[2] += t.x; f[3] += t.y;
              t = unpack_half2(b0.z); f[4] += t.x; f[5] += t.y;   t = unpack_half2(b0.w); f[6] += t.x; f[7] += t.y;
              t = unpack_half2(b1.x); f[8] += t.x; f[9] += t.y;   t = unpack_half2(b1.y); f[10] += t.x; f[11] += t.y;
              t = unpack_half2(b1.z); f[12] += t.x; f[13] += t.y; t = unpack_half2(b1.w); f[14] += t.x; f[15] += t.y;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col + i < p.n_out) f[i] += __half2float(p.rowvec[(size_t)bb * p.ld_rowvec + col + i]);
            }
          }
          // two 16-byte chunks of this row, swizzled (chunk index XOR row phase) inside their 32-column sub-tile
          uint8_t* sub = my_row + (c >> 5) * S::kSubBytes;
          const uint32_t j0 = (uint32_t)((c & 31) >> 3);
          uint4* p0 = reinterpret_cast<uint4*>(sub + ((j0 ^ row_sw) << 4));
          uint4* p1 = reinterpret_cast<uint4*>(sub + (((j0 + 1) ^ row_sw) << 4));
          if (p.has_residual) {
            const uint4 r0 = *p0, r1 = *p1;
            float2 t;
            t = unpack_half2(r0.x); f[0] += t.x; f[1] += t.y;   t = unpack_half2(r0.y); f[2] += t.x; f[3] += t.y;
            t = unpack_half2(r0.z); f[4] += t.x; f[5] += t.y;   t = unpack_half2(r0.w); f[6] += t.x; f[7] += t.y;
            t = unpack_half2(r1.x); f[8] += t.x; f[9] += t.y;   t = unpack_half2(r1.y); f[10] += t.x; f[11] += t.y;
            t = unpack_half2(r1.z); f[12] += t.x; f[13] += t.y; t = unpack_half2(r1.w); f[14] += t.x; f[15] += t.y;
          }
          uint4 o0, o1;
          o0.x = cvt_pack_half2(f[0], f[1]);   o0.y = cvt_pack_half2(f[2], f[3]);
          o0.z = cvt_pack_half2(f[4], f[5]);   o0.w = cvt_pack_half2(f[6], f[7]);
          o1.x = cvt_pack_half2(f[8], f[9]);   o1.y = cvt_pack_half2(f[10], f[11]);
          o1.z = cvt_pack_half2(f[12], f[13]); o1.w = cvt_pack_half2(f[14], f[15]);
          *p0 = o0; *p1 = o1;
        }
      } else {
        // GEGLU: tile columns [0, N/2) = value, [N/2, N) = gate, same output channels; this thread owns 32 outputs.
        constexpr int kHalf = kBlockN / 2;
        constexpr int kOutPerThread = kHalf / 2;
        static_assert(kOutPerThread == 32, "GEGLU epilogue assumes kBlockN == 128");
        uint32_t va[32], vg[32];
        tmem_ld32(t_row + hf * kOutPerThread, va);
        tmem_ld32(t_row + kHalf + hf * kOutPerThread, vg);
        tmem_ld_wait();
        uint8_t* sub = my_row + hf * S::kSubBytes;   // 32 outputs == one sub-tile
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = hf * kOutPerThread + jj * 8 + i;
            const float a = __uint_as_float(va[jj * 8 + i]) + sBias[c];
            const float g = __uint_as_float(vg[jj * 8 + i]) + sBias[kHalf + c];
            o[i] = a * gelu_erf_fast(g);
          }
          uint4 w;
          w.x = cvt_pack_half2(o[0], o[1]); w.y = cvt_pack_half2(o[2], o[3]);
          w.z = cvt_pack_half2(o[4], o[5]); w.w = cvt_pack_half2(o[6], o[7]);
          *reinterpret_cast<uint4*>(sub + (((uint32_t)jj ^ row_sw) << 4)) = w;
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      fence_proxy_async();                              // generic-proxy smem writes -> visible to the TMA store
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
#pragma unroll
        for (int s = 0; s < S::kSubTiles; ++s) {
          const int col = nt * S::kOutCols + s * 32;
          if (col < p.n_out) tma_store_4d(&mapO, sOut + s * S::kSubBytes, col, tc.x0, tc.y0, tc.b0);
        }
        tma_store_commit();
        tma_store_wait_read();       // the staging tile may be overwritten again
        const int next = tile + gridDim.x;
        if (p.has_residual && next < total_tiles) load_residual(next);
      }
    }
    if (et == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

}  // namespace dg
