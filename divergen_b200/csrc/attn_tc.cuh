// tcgen05 flash-style attention: out = softmax(Q K^T * d^-1/2) V for one (batch, head) and 256 query rows per CTA.
//
// Two 128-row query tiles per CTA ping-pong on the tensor pipe (each has its own S/P and O accumulators in TMEM and
// its own 4-warp softmax group), sharing every K/V tile that TMA streams through a multi-stage smem ring:
//   S_q   = Q_q K_j^T           tcgen05.mma SS (both operands K-major, 128-byte swizzle)      -> TMEM fp32
//   P_q   = exp2(S_q*c - m)     softmax warps: tcgen05.ld -> registers -> fp16 pairs -> tcgen05.st (P aliases S)
//   O_q  += P_q V_j             tcgen05.mma TS (A = P from TMEM, B = V tile, MN-major)        -> TMEM fp32
// With kSBuf = 2 every query tile owns TWO score buffers: S_q(j+2) is issued right after P_q(j) V_j, so the scores of the
// next key tile are already in TMEM when the softmax group finishes the current one -- the softmax groups (MUFU / issue
// bound) run back to back and the tensor pipe works in their shadow instead of in series with them.
// Online softmax keeps a (lazily updated) running max; O is rescaled in TMEM only when the max grows by > 2^8.
// The S x S score matrix never exists in HBM.  Head dims that are not multiples of 64 (SD-1.5: 40/80/160) are
// zero-padded by TMA out-of-bounds fill, keys beyond Sk (cross-attention: 77) are masked to -inf.
//
// Replaces (reference side): diffusers Attention / F.scaled_dot_product_attention / xformers inside
// BasicTransformerBlock.attn1/attn2, reached from DiverGen/generation/txt2img_diffusers_stages_from_txt.py:255-259
// (xformers enabled at :186).
#pragma once
#include "common.cuh"

namespace dg {

// fp32 -> fp16 conversion of the softmax weights WITHOUT the conversion instruction: F2FP shares the 16-lane XU pipe with
// MUFU.EX2, and together they are this kernel's roofline (measured: XU 77 % busy, 149 XU instructions per 128 keys per
// row-warp, 64 of them conversions).  The weights are produced pre-scaled by 2^-kPShift = 2^-(112 - 9):
//   v = 2^(x - m - kPShift)  has the fp32 exponent field of an fp16 number  =>  fp16 bits of P' = 2^9 * P are bits(v) >> 13.
// P' <= 2^9 * 2^6 (lazy-max slack) = 2^15 stays finite in fp16; v < 2^-126 (P < 2^-23, the fp16 subnormal floor of the
// unscaled weights) is flushed to zero by ex2.ftz.  Truncation instead of round-to-nearest biases every weight by
// E[-2^-11 / mantissa] = -3.52e-4 relative; the row sum l is taken over the unrounded v, so 1/l is corrected by that factor.
// Measured (round 1): 8 % SLOWER than F2FP with 8 softmax warps per SM (the three extra ALU instructions per pair cost
// more issue slots than the XU time they free), so it is off; kept for the 16-softmax-warp variant's A/B.
#ifndef DG_ATTN_BITPACK
#define DG_ATTN_BITPACK 0
#endif
constexpr bool kBitPackP = DG_ATTN_BITPACK != 0;
constexpr float kPShift = kBitPackP ? 103.0f : 0.0f;
constexpr float kLazyMax = kBitPackP ? 6.0f : 8.0f;
__device__ __forceinline__ uint32_t pack_p_bits(float v0, float v1) {
  if constexpr (kBitPackP) return (__float_as_uint(v0) >> 13) | ((__float_as_uint(v1) << 3) & 0xFFFF0000u);
  else return cvt_pack_half2(v0, v1);
}

struct AttnParams {
  int Sq, Sk;        // query / key tokens per sample
  int heads;
  int ldo;           // output row stride (elements) = heads*d
  float scale_log2;  // d^-1/2 * log2(e)
  int causal;        // 1: key t is visible to query row r only if t <= r (CLIP text encoder); kSplit == 1 variants only
  __half* out;       // [B, Sq, heads*d]
  // Wave-balanced grid (plan_attn_grid, host): 1-D grid; the (batch, head) pairs form two groups -- the first n_g1 pairs are cut
  // into a1 CTAs of kQTiles query tiles + b1 CTAs of kQTiles - 1, the others into a2 + b2 -- and every full-size CTA comes
  // before every short one in block order (longest first), so the block scheduler packs the SMs to within one tile:
  // 64 x 32 tiles on 148 one-CTA SMs take 14 tile times instead of 4 waves x 4 tiles.  part_on = 0: grid (x, head, batch).
  int part_on, n_bh, n_g1, a1, b1, a2, b2;
};

// kQ = softmax streams per CTA (one 4-warp group, one score buffer set and one O accumulator each).  kSplit = 1: every
// stream is its own 128-row query tile.  kSplit = 2: two streams share a query tile and take the even / odd key tiles
// (flash-decoding style); their (m, l, O) are merged through shared memory at the end -- 16 softmax warps per SM with
// 256-row CTAs, so the 64x64-level grid is 1024 CTAs (6.9 waves) instead of 512 (3.46 waves of 512-row CTAs).
template <int kD, int kKV, int kStages, int kSBuf = 1, int kQ = 2, int kSplit = 1>
struct AttnCfg {
  static constexpr int kQTiles = kQ / kSplit;           // 128-row query tiles per CTA
  static constexpr int kChunks = (kD + 63) / 64;        // 64-wide (128 B) head-dim chunks
  static constexpr int kDPad = (kD + 15) / 16 * 16;     // MMA-K of QK^T and MMA-N of PV
  static constexpr int kQTileBytes = kChunks * 128 * 128;
  static constexpr int kKVChunkBytes = kKV * 128;
  static constexpr int kKTileBytes = kChunks * kKVChunkBytes;
  static constexpr int kStageBytes = 2 * kKTileBytes;   // K then V
  static constexpr int kExFloats = (kSplit > 1) ? kQTiles * (kDPad + 2) * 128 : 0;   // (m, l, O[kDPad]) x 128 rows, column-major
  static constexpr int kSmem = kQTiles * kQTileBytes + kStages * kStageBytes + kExFloats * 4 + 1024 + 512;
  static constexpr int kThreads = 128 + kQ * 128;       // producer / MMA / TMEM / spare warp + one 4-warp softmax group per query tile
  // TMEM columns
  static constexpr int kSStride = kSBuf * kKV;          // score buffer b of query tile q: q * kSStride + b * kKV
  static constexpr int kO0 = kQ * kSStride;
  static constexpr int kOStride = (kQ > 2 && kDPad <= 64) ? 64 : (kDPad <= 128) ? 128 : 192;
  static_assert(kO0 + (kQ - 1) * kOStride + kDPad <= 512, "TMEM budget");
  static_assert(kStages > kSBuf * kSplit, "the K tile of S(j + kSBuf * kSplit) and the V tile of PV(j) are live together");
  static_assert(kSplit == 1 || kSplit == 2, "one or two key-tile streams per query tile");
};

template <int kD, int kKV, int kStages, int kPoly, int kSBuf, int kQ, int kSplit>   // kPoly: every kPoly-th pair of softmax elements takes the polynomial exp2 (0 = none)
__global__ void __launch_bounds__(128 + kQ * 128, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapV, const AttnParams p) {
  using C = AttnCfg<kD, kKV, kStages, kSBuf, kQ, kSplit>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + C::kQTiles * C::kQTileBytes;
  float* sEx = reinterpret_cast<float*>(sKV + kStages * C::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kStages * C::kStageBytes + C::kExFloats * 4);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* kv_full = bars + 1;            // [kStages]
  uint64_t* kv_empty = kv_full + kStages;  // [kStages]
  uint64_t* s_full = kv_empty + kStages;   // [kQ][kSBuf]
  uint64_t* p_full = s_full + kQ * kSBuf;  // [kQ][kSBuf]  (a softmax warp may run one key tile ahead of its group: per-buffer barriers)
  uint64_t* o_done = p_full + kQ * kSBuf;  // [kQ]  last PV of the tile
  uint64_t* pv_done = o_done + kQ;         // [kQ]  every PV (the rare O rescale must not race the accumulate in flight)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + kQ);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int q_row0 = blockIdx.x * (C::kQTiles * 128);
  int head = blockIdx.y;
  int batch = blockIdx.z;
  int nq = C::kQTiles;                     // query tiles this CTA works on (the streams of the others stay idle)
  if (p.part_on) {
    int idx = blockIdx.x, bh, tile0;
    const int f1 = p.n_g1 * p.a1, f2 = (p.n_bh - p.n_g1) * p.a2, t1 = p.n_g1 * p.b1;
    if (idx < f1) { bh = idx / p.a1; tile0 = (idx % p.a1) * C::kQTiles; }
    else if (idx < f1 + f2) { idx -= f1; bh = p.n_g1 + idx / p.a2; tile0 = (idx % p.a2) * C::kQTiles; }
    else if (idx < f1 + f2 + t1) { idx -= f1 + f2; bh = idx / p.b1; tile0 = p.a1 * C::kQTiles + (idx % p.b1) * (C::kQTiles - 1); nq = C::kQTiles - 1; }
    else { idx -= f1 + f2 + t1; bh = p.n_g1 + idx / p.b2; tile0 = p.a2 * C::kQTiles + (idx % p.b2) * (C::kQTiles - 1); nq = C::kQTiles - 1; }
    q_row0 = tile0 * 128; head = bh % p.heads; batch = bh / p.heads;
  }
  const int nkv = (p.Sk + kKV - 1) / kKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < kQ * kSBuf; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    for (int i = 0; i < kQ; ++i) { mbar_init(&o_done[i], 1); mbar_init(&pv_done[i], 1); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, C::kQTiles * C::kQTileBytes);
      for (int q = 0; q < C::kQTiles; ++q)
        for (int c = 0; c < C::kChunks; ++c)
          tma_load_4d(sQ + q * C::kQTileBytes + c * (128 * 128), &mapQ, q_full, c * 64, head, q_row0 + q * 128, batch);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait_parked(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + stage * C::kStageBytes;
        uint8_t* sv = sk + C::kKTileBytes;
        mbar_arrive_expect_tx(&kv_full[stage], C::kStageBytes);
        for (int c = 0; c < C::kChunks; ++c) {
          tma_load_4d(sk + c * C::kKVChunkBytes, &mapK, &kv_full[stage], c * 64, head, j * kKV, batch);
          tma_load_4d(sv + c * C::kKVChunkBytes, &mapV, &kv_full[stage], c * 64, head, j * kKV, batch);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_f16(kKV, false);
    constexpr uint32_t idesc_o = make_idesc_f16(C::kDPad, true);
    const uint32_t sq_addr = smem_u32(sQ);

    auto issue_S = [&](int q, int b, uint32_t sk_addr) {
#pragma unroll
      for (int s = 0; s < C::kDPad / 16; ++s) {
        const int c = s >> 2, k = s & 3;
        const uint64_t da = make_smem_desc_sw128(sq_addr + (q / kSplit) * C::kQTileBytes + c * (128 * 128) + k * 32, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(sk_addr + c * C::kKVChunkBytes + k * 32, 16, 1024);
        umma_ss(tmem_base + q * C::kSStride + b * kKV, da, db, idesc_s, s ? 1u : 0u);
      }
      umma_commit(&s_full[q * kSBuf + b]);
    };
    auto issue_PV = [&](int q, int b, uint32_t sv_addr, bool accum) {
#pragma unroll
      for (int s = 0; s < kKV / 16; ++s) {
        // V tile: [keys][64-wide d chunk] rows of 128 B => MN-major; 16 keys = 2048 B along K.
        const uint64_t db = make_smem_desc_sw128(sv_addr + s * 2048, C::kKVChunkBytes, 1024);
        umma_ts(tmem_base + C::kO0 + q * C::kOStride, tmem_base + q * C::kSStride + b * kKV + s * 8, db, idesc_o, (accum || s) ? 1u : 0u);
      }
    };

    mbar_wait_parked(q_full, 0);
    // stream s = qq * kSplit + h works on the key tiles j with j % kSplit == h; its local iteration is jj = j / kSplit and
    // its score buffer b = jj % kSBuf.  After P_s(j) V_j the scores of its tile j + kSplit * kSBuf are issued into buffer b.
    int ready = 0;                           // key tiles whose K/V have landed (kv_full waited)
    auto ensure = [&](int t) {
      while (ready <= t) { mbar_wait_parked(&kv_full[ready % kStages], (ready / kStages) & 1); ++ready; }
      tc_fence_after();
    };
    constexpr int kAhead = kSplit * kSBuf;
    for (int t = 0; t < kAhead && t < nkv; ++t) {
      ensure(t);
      if (elect_one()) {
        for (int qq = 0; qq < nq; ++qq)
          issue_S(qq * kSplit + t % kSplit, (t / kSplit) % kSBuf, smem_u32(sKV + (t % kStages) * C::kStageBytes));
      }
      __syncwarp();
    }
    for (int j = 0; j < nkv; ++j) {
      const int stage = j % kStages;
      const uint32_t sv_addr = smem_u32(sKV + stage * C::kStageBytes + C::kKTileBytes);
      const int jj = j / kSplit, b = jj % kSBuf;
      const int tn = j + kAhead;             // the tile whose scores go into the buffer this step frees
      const bool has_next = tn < nkv;
      if (has_next) ensure(tn);
      const uint32_t sk_next = smem_u32(sKV + (tn % kStages) * C::kStageBytes);
      const bool last = (j + kSplit >= nkv); // the stream's last tile
      for (int qq = 0; qq < nq; ++qq) {
        const int st = qq * kSplit + j % kSplit;
        mbar_wait_parked(&p_full[st * kSBuf + b], (jj / kSBuf) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_PV(st, b, sv_addr, jj > 0);
          umma_commit(last ? &o_done[st] : &pv_done[st]);
          if (has_next) issue_S(st, b, sk_next);
          if (qq == nq - 1) umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / correction / epilogue =====================
    const int q = (warp - 4) >> 2;       // stream handled by this warp group
    const int qt = q / kSplit;           // its query tile
    const int h = q % kSplit;            // its share of the key tiles: j % kSplit == h
    const bool active = qt < nq;         // (a short CTA of the wave-balanced grid leaves its last query tile's streams idle)
    const int nkv_s = active ? nkv : 0;
    const int quad = warp & 3;           // TMEM lane quadrant
    const int row = quad * 32 + lane;    // row within the 128-row tile
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tS_q = tmem_base + (uint32_t)(q * C::kSStride) + lane_off;
    const uint32_t tO = tmem_base + (uint32_t)(C::kO0 + q * C::kOStride) + lane_off;
    // with 2 softmax groups per SM the TMEM loads are software-pipelined one 32-column chunk ahead of the math; with 4
    // groups (16 warps, <= 96 registers) the other warps of the scheduler hide that latency instead
    constexpr bool kPre = kQ <= 2;
    float m_run = -INFINITY;  // running max (log2 domain, already scaled)
    float l_run = 0.f;
    for (int jj = 0, j = h; j < nkv_s; ++jj, j += kSplit) {
      const uint32_t tS = tS_q + (uint32_t)((jj % kSBuf) * kKV);
      mbar_wait(&s_full[q * kSBuf + jj % kSBuf], (jj / kSBuf) & 1);
      tc_fence_after();
      // keys of this tile visible to this thread's row: the tile's valid keys, cut at the diagonal under a causal mask
      // (a row always sees key 0, so the first tile never masks a whole row)
      const int kv_tile = min(kKV, p.Sk - j * kKV);
      const int kv_valid = p.causal ? min(kv_tile, q_row0 + qt * 128 + row + 1 - j * kKV) : kv_tile;
      const bool full_tile = kv_tile == kKV && !p.causal;
      // pass 1: row max (FMNMX3: two columns per instruction, two independent chains)
      // (TMEM loads are software-pipelined one 32-column chunk ahead of the math in both passes: the single warp that owns
      // these rows would otherwise expose the tcgen05.ld latency eight times per tile)
      float mx = -INFINITY, mx_b = -INFINITY;
      uint32_t vv[kPre ? 2 : 1][32];
      if (kPre) { tmem_ld32(tS, vv[0]); tmem_ld_wait(); }
#pragma unroll
      for (int c = 0; c < kKV; c += 32) {
        uint32_t* v = vv[kPre ? (c >> 5) & 1 : 0];
        if (!kPre) { tmem_ld32(tS + c, v); tmem_ld_wait(); }
        if (kPre && c + 32 < kKV) tmem_ld32(tS + c + 32, vv[kPre ? ((c >> 5) + 1) & 1 : 0]);
        if (full_tile) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx = max3f(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            mx_b = max3f(mx_b, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float s = __uint_as_float(v[i]);
            if (c + i >= kv_valid) s = -INFINITY;
            mx = fmaxf(mx, s);
          }
        }
        if (kPre && c + 32 < kKV) tmem_ld_wait();
      }
      mx = fmaxf(mx, mx_b);
      const float m_tile = mx * p.scale_log2;
      if (jj == 0) {
        m_run = m_tile;
      } else {
        // lazy rescale: only when the max grew enough to threaten the fp16 range of P.  tcgen05.ld/st are warp-wide
        // (.sync.aligned), so the decision is made per warp and rows that do not need it scale by exactly 1.
        const bool need = m_tile > m_run + kLazyMax;
        if (__any_sync(0xffffffffu, need)) {
          // S(j) was issued before PV(j-1): wait for that accumulate to land before touching O
          if (kSBuf > 1) { mbar_wait(&pv_done[q], (jj - 1) & 1); tc_fence_after(); }
          const float alpha = need ? exp2f(m_run - m_tile) : 1.0f;
          if (need) { m_run = m_tile; l_run *= alpha; }
#pragma unroll
          for (int c = 0; c < C::kDPad; c += 16) {
            uint32_t o[16];
            tmem_ld16(tO + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + c, o);
          }
        }
      }
      // pass 2: P = exp2(S*c - m), packed fp16 over the front half of the S columns.
      // Per pair of columns: 1 FFMA2 + 2 MUFU.EX2 + 1 F2FP + 1 FADD2 -- the MUFU pipe (16/clk/SM) is the bound.
      float lsum = 0.f;
      if (full_tile) {
        const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
        const uint64_t nm2 = pack_f32x2(-(m_run + kPShift), -(m_run + kPShift));
        uint64_t sum2 = pack_f32x2(0.f, 0.f);
        if (kPre) { tmem_ld32(tS, vv[0]); tmem_ld_wait(); }
#pragma unroll
        for (int c = 0; c < kKV; c += 32) {
          uint32_t* v = vv[kPre ? (c >> 5) & 1 : 0];
          if (!kPre) { tmem_ld32(tS + c, v); tmem_ld_wait(); }
          // the next chunk's load must not overtake this chunk's store into the same TMEM columns: P (fp16 pairs) of chunk c
          // lands in columns [c/2, c/2+16), which chunk c+32 (columns [c+32, c+64)) never overlaps for c >= 0
          if (kPre && c + 32 < kKV) tmem_ld32(tS + c + 32, vv[kPre ? ((c >> 5) + 1) & 1 : 0]);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nm2);
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            float p0, p1;
            if (kPoly > 0 && ((i >> 1) % (kPoly > 0 ? kPoly : 1)) == (kPoly > 0 ? kPoly : 1) - 1) {
              // every third pair: exp2 on the FMA/ALU pipes (the MUFU pipe, 16 results/clk/SM, is this kernel's roofline).
              // 2^x = 2^n * 2^f, n = round(x) via the 1.5*2^23 magic add, 2^f by a degree-3 minimax polynomial on
              // [-0.5, 0.5] (rel. error 7.7e-5, below the fp16 rounding of P), 2^n by adding n to the exponent field.
              const uint64_t xc = pack_f32x2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
              const uint64_t xr = add_f32x2(xc, pack_f32x2(12582912.0f, 12582912.0f));
              const uint64_t nn = add_f32x2(xr, pack_f32x2(-12582912.0f, -12582912.0f));
              const uint64_t f2 = fma_f32x2(nn, pack_f32x2(-1.0f, -1.0f), xc);
              uint64_t q2 = fma_f32x2(pack_f32x2(0.05508868396282196f, 0.05508868396282196f), f2,
                                      pack_f32x2(0.24260404706001282f, 0.24260404706001282f));
              q2 = fma_f32x2(q2, f2, pack_f32x2(0.6932762265205383f, 0.6932762265205383f));
              q2 = fma_f32x2(q2, f2, pack_f32x2(0.9999289512634277f, 0.9999289512634277f));
              float q0, q1, r0, r1;
              unpack_f32x2(q2, q0, q1);
              unpack_f32x2(xr, r0, r1);
              p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
              p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
            } else {
              p0 = fast_exp2(x0); p1 = fast_exp2(x1);
            }
            pk[i >> 1] = pack_p_bits(p0, p1);
            sum2 = add_f32x2(sum2, pack_f32x2(p0, p1));
          }
          if (kPre && c + 32 < kKV) tmem_ld_wait();     // chunk c+32 is in registers before P(c) overwrites columns it might share
          tmem_st16(tS + (c >> 1), pk);
        }
        float s0, s1;
        unpack_f32x2(sum2, s0, s1);
        lsum = s0 + s1;
      } else {
#pragma unroll
        for (int c = 0; c < kKV; c += 32) {
          uint32_t v[32];
          tmem_ld32(tS + c, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c + i < kv_valid) ? fast_exp2(__uint_as_float(v[i]) * p.scale_log2 - (m_run + kPShift)) : 0.f;
            const float p1 = (c + i + 1 < kv_valid) ? fast_exp2(__uint_as_float(v[i + 1]) * p.scale_log2 - (m_run + kPShift)) : 0.f;
            pk[i >> 1] = pack_p_bits(p0, p1);
            lsum += p0 + p1;
          }
          tmem_st16(tS + (c >> 1), pk);
        }
      }
      l_run += lsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[q * kSBuf + jj % kSBuf]);
    }
    // epilogue: O / l -> fp16 global
    const bool has_tiles = h < nkv_s;    // (a second stream has nothing to do when there is a single key tile)
    if (has_tiles) { mbar_wait(&o_done[q], 0); tc_fence_after(); }
    const int qrow = q_row0 + qt * 128 + row;
    // O = sum(P' V) with P' = 2^112 v (truncated), l = sum(v):  out = O * 2^-112 / l, corrected for the truncation bias
    constexpr float kInvScale = kBitPackP ? 1.9259299e-34f * 1.000352f : 1.0f;
    float a_own = 1.0f, a_oth = 0.0f;    // merge weights of this stream's and the partner stream's accumulators
    const float* ex = sEx + qt * (C::kDPad + 2) * 128 + row;   // partner's (m, l, O[..]) of this row, element e at ex[e * 128]
    if constexpr (kSplit == 2) {
      if (h == 1) {
        float* exw = sEx + qt * (C::kDPad + 2) * 128 + row;
        exw[0] = has_tiles ? m_run : -INFINITY;
        exw[128] = has_tiles ? l_run : 0.f;
#pragma unroll
        for (int c = 0; c < C::kDPad; c += 16) {
          uint32_t o[16];
          if (has_tiles) { tmem_ld16(tO + c, o); tmem_ld_wait(); }
#pragma unroll
          for (int i = 0; i < 16; ++i) exw[(2 + c + i) * 128] = has_tiles ? __uint_as_float(o[i]) : 0.f;
        }
        asm volatile("bar.arrive %0, 256;" ::"r"(1 + qt) : "memory");
      } else {
        asm volatile("bar.sync %0, 256;" ::"r"(1 + qt) : "memory");
        const float m1 = ex[0], l1 = ex[128];
        const float m = fmaxf(m_run, m1);
        a_own = exp2f(m_run - m);
        a_oth = (l1 > 0.f) ? exp2f(m1 - m) : 0.f;
        l_run = l_run * a_own + l1 * a_oth;
      }
    }
    const float inv_l = (kSplit == 2 && h == 1) ? 0.f : __fdividef(kInvScale, l_run);
    __half* orow = p.out + ((size_t)batch * p.Sq + qrow) * p.ldo + head * kD;
#pragma unroll
    for (int c = 0; c < C::kDPad; c += 16) {
      if (kSplit == 2 && h == 1) break;  // the partner stream writes the merged rows
      if (!active) break;
      uint32_t o[16];
      tmem_ld16(tO + c, o);
      tmem_ld_wait();
      if constexpr (kSplit == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          o[i] = __float_as_uint(fmaf(__uint_as_float(o[i]), a_own, ex[(2 + c + i) * 128] * a_oth));
      }
      if (qrow < p.Sq) {
        uint4 w0, w1;
        w0.x = pack_half2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l);
        w0.y = pack_half2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l);
        w0.z = pack_half2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l);
        w0.w = pack_half2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l);
        w1.x = pack_half2(__uint_as_float(o[8]) * inv_l, __uint_as_float(o[9]) * inv_l);
        w1.y = pack_half2(__uint_as_float(o[10]) * inv_l, __uint_as_float(o[11]) * inv_l);
        w1.z = pack_half2(__uint_as_float(o[12]) * inv_l, __uint_as_float(o[13]) * inv_l);
        w1.w = pack_half2(__uint_as_float(o[14]) * inv_l, __uint_as_float(o[15]) * inv_l);
        if (c + 8 <= kD) *reinterpret_cast<uint4*>(orow + c) = w0;
        if (c + 16 <= kD) *reinterpret_cast<uint4*>(orow + c + 8) = w1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

}  // namespace dg
