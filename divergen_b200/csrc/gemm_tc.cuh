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
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4..7 = epilogue (TMEM -> registers -> smem staging -> coalesced fp16 global).  Persistent over tiles, kStages-deep smem ring, two TMEM
// accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
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
  int n_gemm;       // GEMM N (rows of Wt that are meaningful); = 2*ceil-padded n_out for GEGLU
  int ldo;          // output row stride in elements
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
  const __half* residual;  // [pixels, ld_res] or nullptr
  int ld_res;
  int geglu;               // 1: tile = [a(half) | g(half)] -> a*gelu(g)
  __half* out;
  float* out_f32;          // optional fp32 output (split-K partials not used yet; kept null)
};

template <int kBlockN, int kStages>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBlockN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutRowBytes = kBlockN * 2 + 16;  // +16 B pad: conflict-free 16-byte row-strided stores
  static constexpr int kOutBytes = kBlockM * kOutRowBytes + kBlockM * 8 /*row -> pixel table*/;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kOutBytes;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int kBlockN, int kStages>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapW, const GemmParams p) {
  using S = GemmSmem<kBlockN, kStages>;
  static_assert(kBlockN % 16 == 0 && kBlockN <= 256, "UMMA N");
  static_assert(S::kBBytes % 1024 == 0, "B stage must keep 1024-byte alignment");
  constexpr uint32_t kAccStride = 256;  // TMEM columns between the two accumulator stages
  constexpr uint32_t kTmemCols = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
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
        const int nt = tile % p.tiles_n;
        int mt = tile / p.tiles_n;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bn;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dx = (p.taps == 9) ? (tap % 3 - 1) : 0;
          const int dy = (p.taps == 9) ? (tap / 3 - 1) : 0;
          for (int kb = 0; kb < kb_per_tap; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_arrive_expect_tx(&full[stage], S::kStageBytes);
            if (kb < p.kb0) tma_load_4d(sa, &mapA0, &full[stage], kb * kBlockK, x0 + dx, y0 + dy, b0);
            else            tma_load_4d(sa, &mapA1, &full[stage], (kb - p.kb0) * kBlockK, x0 + dx, y0 + dy, b0);
            tma_load_2d(sb, &mapW, &full[stage], (tap * kb_per_tap + kb) * kBlockK, nt * kBlockN);
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
    // ===================== epilogue =====================
    // pass 1 (thread = accumulator row): TMEM -> registers -> +bias (+time-embedding vector) -> fp16 -> padded smem tile;
    //         the TMEM stage is released as soon as it has been drained.
    // pass 2 (coalesced): the 128 threads sweep the staged tile in 16-byte vectors, add the residual (independent,
    //         batched global loads) and store full 32-byte sectors.
    const int q = warp - 4;          // TMEM lane quadrant
    const int r = q * 32 + lane;     // row within the tile
    const int et = threadIdx.x - 128;  // 0..127
    const int box_xy = p.bw * p.bh;
    uint8_t* sOut = smem + kStages * S::kStageBytes + 256;
    long long* sRow = reinterpret_cast<long long*>(sOut + kBlockM * S::kOutRowBytes);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n;
      int mt = tile / p.tiles_n;
      const int tx = mt % p.tiles_x; mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int tb = mt / p.tiles_y;
      const int xx = tx * p.bw + r % p.bw;
      const int yy = ty * p.bh + (r / p.bw) % p.bh;
      const int bb = tb * p.bn + r / box_xy;
      const bool row_ok = (xx < p.W) && (yy < p.H) && (bb < p.B);
      const long long pix = ((long long)bb * p.H + yy) * p.W + xx;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);
      // previous tile's pass 2 must be done with the staging tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
      sRow[r] = row_ok ? pix : -1;
      uint8_t* my_row = sOut + r * S::kOutRowBytes;

      if (!p.geglu) {
        const int n0 = nt * kBlockN;
#pragma unroll 2
        for (int c = 0; c < kBlockN; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_row + c, v);
          tmem_ld_wait();
          const int col = n0 + c;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (col + 16 <= p.n_gemm) {
            if (p.bias) {
              const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col);
              const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
              float2 t;
              t = unpack_half2(b0.x); f[0] += t.x; f[1] += t.y;   t = unpack_half2(b0.y); f[2] += t.x; f[3] += t.y;
              t = unpack_half2(b0.z); f[4] += t.x; f[5] += t.y;   t = unpack_half2(b0.w); f[6] += t.x; f[7] += t.y;
              t = unpack_half2(b1.x); f[8] += t.x; f[9] += t.y;   t = unpack_half2(b1.y); f[10] += t.x; f[11] += t.y;
              t = unpack_half2(b1.z); f[12] += t.x; f[13] += t.y; t = unpack_half2(b1.w); f[14] += t.x; f[15] += t.y;
            }
            if (p.rowvec && row_ok) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.rowvec + (size_t)bb * p.ld_rowvec + col);
              const uint4 b0 = __ldg(rp), b1 = __ldg(rp + 1);
              float2 t;
              t = unpack_half2(b0.x); f[0] += t.x; f[1] += t.y;   t = unpack_half2(b0.y); f[2] += t.x; f[3] += t.y;
              t = unpack_half2(b0.z); f[4] += t.x; f[5] += t.y;   t = unpack_half2(b0.w); f[6] += t.x; f[7] += t.y;
              t = unpack_half2(b1.x); f[8] += t.x; f[9] += t.y;   t = unpack_half2(b1.y); f[10] += t.x; f[11] += t.y;
              t = unpack_half2(b1.z); f[12] += t.x; f[13] += t.y; t = unpack_half2(b1.w); f[14] += t.x; f[15] += t.y;
            }
          } else {
            // ragged N tail (e.g. conv_out: 4 channels): guarded scalar loads
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (col + i < p.n_gemm) {
                if (p.bias) f[i] += __half2float(p.bias[col + i]);
                if (p.rowvec && row_ok) f[i] += __half2float(p.rowvec[(size_t)bb * p.ld_rowvec + col + i]);
              }
            }
          }
          uint4 o0, o1;
          o0.x = pack_half2(f[0], f[1]);   o0.y = pack_half2(f[2], f[3]);
          o0.z = pack_half2(f[4], f[5]);   o0.w = pack_half2(f[6], f[7]);
          o1.x = pack_half2(f[8], f[9]);   o1.y = pack_half2(f[10], f[11]);
          o1.z = pack_half2(f[12], f[13]); o1.w = pack_half2(f[14], f[15]);
          uint4* sp = reinterpret_cast<uint4*>(my_row + c * 2);
          sp[0] = o0; sp[1] = o1;
        }
      } else {
        // GEGLU: columns [0, kBlockN/2) = value, [kBlockN/2, kBlockN) = gate, same output channels.
        constexpr int kHalf = kBlockN / 2;
#pragma unroll 1
        for (int c = 0; c < kHalf; c += 16) {
          uint32_t va[16], vg[16];
          tmem_ld16(t_row + c, va);
          tmem_ld16(t_row + kHalf + c, vg);
          tmem_ld_wait();
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = __uint_as_float(va[i]) + __half2float(__ldg(p.bias + nt * kBlockN + c + i));
            const float g = __uint_as_float(vg[i]) + __half2float(__ldg(p.bias + nt * kBlockN + kHalf + c + i));
            o[i] = a * gelu_erf(g);
          }
          uint4 o0, o1;
          o0.x = pack_half2(o[0], o[1]);   o0.y = pack_half2(o[2], o[3]);
          o0.z = pack_half2(o[4], o[5]);   o0.w = pack_half2(o[6], o[7]);
          o1.x = pack_half2(o[8], o[9]);   o1.y = pack_half2(o[10], o[11]);
          o1.z = pack_half2(o[12], o[13]); o1.w = pack_half2(o[14], o[15]);
          uint4* sp = reinterpret_cast<uint4*>(my_row + c * 2);
          sp[0] = o0; sp[1] = o1;
        }
      }
      // accumulator drained: hand the TMEM stage back to the MMA warp before the global-memory pass
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      asm volatile("bar.sync 1, 128;" ::: "memory");

      // ---- pass 2
      const int tile_cols = p.geglu ? kBlockN / 2 : kBlockN;
      const int n0 = nt * tile_cols;
      const int vec_per_row = tile_cols / 8;
      const int total_vec = kBlockM * vec_per_row;
      constexpr int kBatch = 5;
      for (int base = et; base < total_vec; base += 128 * kBatch) {
        uint4 val[kBatch], res[kBatch];
        long long gofs[kBatch];
        int colv[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          const int v = base + b * 128;
          gofs[b] = -1; colv[b] = 0;
          res[b] = make_uint4(0u, 0u, 0u, 0u);
          if (v < total_vec) {
            const int row = v / vec_per_row, cv = v - row * vec_per_row;
            const long long px = sRow[row];
            const int col = n0 + cv * 8;
            if (px >= 0 && col < p.n_out) {
              gofs[b] = px; colv[b] = col;
              val[b] = *reinterpret_cast<const uint4*>(sOut + row * S::kOutRowBytes + cv * 16);
              if (p.residual && col + 8 <= p.n_out) res[b] = *reinterpret_cast<const uint4*>(p.residual + px * p.ld_res + col);
            }
          }
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
          if (gofs[b] < 0) continue;
          const int col = colv[b];
          if (col + 8 <= p.n_out) {
            uint4 o = val[b];
            if (p.residual) {
              const float2 a0 = unpack_half2(o.x), a1 = unpack_half2(o.y), a2 = unpack_half2(o.z), a3 = unpack_half2(o.w);
              const float2 r0 = unpack_half2(res[b].x), r1 = unpack_half2(res[b].y), r2 = unpack_half2(res[b].z), r3 = unpack_half2(res[b].w);
              o.x = pack_half2(a0.x + r0.x, a0.y + r0.y); o.y = pack_half2(a1.x + r1.x, a1.y + r1.y);
              o.z = pack_half2(a2.x + r2.x, a2.y + r2.y); o.w = pack_half2(a3.x + r3.x, a3.y + r3.y);
            }
            *reinterpret_cast<uint4*>(p.out + gofs[b] * p.ldo + col) = o;
          } else {
            const __half* hv = reinterpret_cast<const __half*>(&val[b]);
            for (int i = 0; i < 8 && col + i < p.n_out; ++i) {
              float x = __half2float(hv[i]);
              if (p.residual) x += __half2float(p.residual[gofs[b] * p.ld_res + col + i]);
              p.out[gofs[b] * p.ldo + col + i] = __float2half_rn(x);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<kTmemCols>(tmem_base); }
}

}  // namespace dg
