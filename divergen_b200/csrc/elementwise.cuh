// HBM-bound kernels of the path: layout changes, GroupNorm / LayerNorm, timestep embedding + small-batch GEMV,
// nearest-neighbour upsample, im2col gathers for the few convolutions TMA cannot tile (4-channel conv_in, stride-2
// downsamplers), and the fused classifier-free-guidance + DDIM step.  All are coalesced, 16-byte vectorised, fp32 math.
#pragma once
#include "common.cuh"

namespace dg {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

struct alignas(16) Half8 { __half2 h[4]; };

__device__ __forceinline__ void load8(const __half* p, float* f) {
  Half8 v = *reinterpret_cast<const Half8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __half22float2(v.h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void store8(__half* p, const float* f) {
  Half8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<Half8*>(p) = v;
}

// ------------------------------------------------------------------ layout
// NCHW fp16 -> NHWC fp16 (tiny tensors only: latents in, noise prediction out).
__global__ void nchw_to_nhwc_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int C, int HW) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n = (size_t)B * C * HW;
  if (i >= n) return;
  int c = i % C; size_t t = i / C; int p = t % HW; int b = t / HW;
  out[i] = in[((size_t)b * C + c) * HW + p];
}
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int C, int HW,
                                    int in_pitch) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n = (size_t)B * C * HW;
  if (i >= n) return;
  int p = i % HW; size_t t = i / HW; int c = t % C; int b = t / C;
  out[i] = in[((size_t)b * HW + p) * in_pitch + c];
}

// Output path: `pt_to_pil`'s arithmetic on the device -- [B,3,H,W] fp16 in [-1,1] -> uint8 [B,H,W,3].
// Mirrors diffusers.utils.pt_to_pil step by step: (x / 2 + 0.5) in fp16, clamp(0, 1), * 255 in fp32, round half to even.
// The host then only PNG-encodes: 786 KB per 512x512 image cross PCIe instead of 1.5 MB, and no float math on a host core.
__global__ void image_to_uint8_hwc_kernel(const __half* __restrict__ img, unsigned char* __restrict__ out, int B, int C, int HW) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;     // one pixel
  if (i >= (size_t)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i % HW);
  for (int c = 0; c < C; ++c) {
    const __half y = __hadd(__hmul(img[((size_t)b * C + c) * HW + p], __float2half_rn(0.5f)), __float2half_rn(0.5f));
    const float v = fminf(fmaxf(__half2float(y), 0.f), 1.f) * 255.0f;
    out[i * C + c] = (unsigned char)__float2int_rn(v);
  }
}

// ------------------------------------------------------------------ GroupNorm (NHWC, optional 2-source concat)
// Work decomposition shared by both kernels: a block owns a strip of `pix_per_block` pixels of one sample; a thread owns
// one 8-channel (16-byte) vector position and walks the strip, so every global access is a coalesced 16-byte load of a
// contiguous NHWC row segment and per-channel constants are computed once per thread, not once per element.
struct GnThreadMap {
  int v0, vstep, pofs, pstride;
  __device__ GnThreadMap(int nvec) {
    if (nvec <= (int)blockDim.x) {
      pstride = blockDim.x / nvec;
      v0 = threadIdx.x % nvec; vstep = nvec;
      pofs = threadIdx.x / nvec;
      if (pofs >= pstride) { v0 = nvec; }  // idle tail threads
    } else {
      pstride = 1; pofs = 0; v0 = threadIdx.x; vstep = blockDim.x;
    }
  }
};

// stats[b][g] = {sum, sumsq} (fp32, atomically accumulated over strips; zeroed by the launcher).
__global__ void gn_stats_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                int groups, int pix_per_block, float* __restrict__ stats) {
  extern __shared__ float sh[];  // [pstride][C][2] partials, then [C][2] channel sums
  const int C = C0 + C1, cpg = C / groups, nvec = C / 8;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const GnThreadMap tm(nvec);
  for (int v = tm.v0; v < nvec; v += tm.vstep) {
    float s[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
    const int c = v * 8;
    const __half* src = (c < C0) ? x0 + c : x1 + (c - C0);
    const int ld = (c < C0) ? C0 : C1;
#pragma unroll 4
    for (int p = p0 + tm.pofs; p < p1; p += tm.pstride) {
      float f[8];
      load8(src + ((size_t)b * HW + p) * ld, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] = fmaf(f[i], f[i], ss[i]); }
    }
    float* dst = sh + ((size_t)tm.pofs * C + c) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[2 * i] = s[i]; dst[2 * i + 1] = ss[i]; }
  }
  __syncthreads();
  // per-channel totals over the pixel sub-strips (into slab 0)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int j = 0; j < tm.pstride; ++j) { a += sh[((size_t)j * C + c) * 2]; q += sh[((size_t)j * C + c) * 2 + 1]; }
    sh[(size_t)c * 2] = a; sh[(size_t)c * 2 + 1] = q;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += sh[(size_t)c * 2]; q += sh[(size_t)c * 2 + 1]; }
    atomicAdd(&stats[((size_t)b * groups + g) * 2], a);
    atomicAdd(&stats[((size_t)b * groups + g) * 2 + 1], q);
  }
}

// y = act((x - mean) * rstd * gamma + beta) = act(x * scale + shift), written as one concatenated NHWC tensor.
__global__ void gn_apply_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                int groups, float eps, int pix_per_block, const float* __restrict__ stats,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta, int do_silu,
                                __half* __restrict__ out) {
  const int C = C0 + C1, cpg = C / groups, nvec = C / 8;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const float inv_n = 1.0f / ((float)cpg * (float)HW);
  const GnThreadMap tm(nvec);
  griddep_launch();
  griddep_wait();            // (returns at once when the kernel was not launched as a programmatic dependent)
  for (int v = tm.v0; v < nvec; v += tm.vstep) {
    const int c = v * 8;
    float sc[8], sf[8], g[8], be[8];
    load8(gamma + c, g);
    load8(beta + c, be);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int grp = (c + k) / cpg;
      const float mean = stats[((size_t)b * groups + grp) * 2] * inv_n;
      const float var = fmaxf(stats[((size_t)b * groups + grp) * 2 + 1] * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      sc[k] = rstd * g[k];
      sf[k] = be[k] - mean * sc[k];
    }
    const __half* src = (c < C0) ? x0 + c : x1 + (c - C0);
    const int ld = (c < C0) ? C0 : C1;
#pragma unroll 4
    for (int p = p0 + tm.pofs; p < p1; p += tm.pstride) {
      const size_t pix = (size_t)b * HW + p;
      float f[8];
      load8(src + pix * ld, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float y = fmaf(f[k], sc[k], sf[k]);
        f[k] = do_silu ? __fdividef(y, 1.0f + __expf(-y)) : y;
      }
      store8(out + pix * C + c, f);
    }
  }
}

#ifndef DG_GN_FOLD_UNROLL
#define DG_GN_FOLD_UNROLL 16
#endif
constexpr int kGnFoldUnroll = DG_GN_FOLD_UNROLL;
// Same apply pass, statistics taken from the producers' fused epilogue sums: stats{0,1}[b][slots][C{0,1}/blk] float2 hold
// (sum, sum of squares) over blk-channel blocks of each 32-pixel slab of each source (gemm2_tc.cuh).  A group of the
// concatenated tensor is a union of whole blocks; each CTA first folds slabs x blocks into 32 (mean, rstd) pairs in shared
// memory, in a fixed order (deterministic).
__global__ void gn_apply_blk_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                    int groups, float eps, int pix_per_block, const float* __restrict__ stats0,
                                    const float* __restrict__ stats1, int blk, int slots, const __half* __restrict__ gamma,
                                    const __half* __restrict__ beta, int do_silu, __half* __restrict__ out) {
  __shared__ float2 s_col[8][256];       // [slab stripe][block column] partial sums
  __shared__ float s_mean[64], s_rstd[64];
  griddep_launch();
  const int C = C0 + C1, cpg = C / groups, nvec = C / 8;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const float inv_n = 1.0f / ((float)cpg * (float)HW);
  // Latency chain of a short launch: statistics -> (barrier) -> affine -> activation loads -> stores.  The affine parameters
  // are weights (no dependence on the producer kernel: loaded before the programmatic-dependency wait) and the first
  // activation vector does not depend on the statistics: both are in flight while the statistics are folded.
  const GnThreadMap tm(nvec);
  const bool have0 = tm.v0 < nvec;
  float g0[8], be0[8], f0[8];
  if (have0) { load8(gamma + tm.v0 * 8, g0); load8(beta + tm.v0 * 8, be0); }
  griddep_wait();
  const int pf = p0 + tm.pofs;
  const bool havep = have0 && pf < p1;
  if (havep) {
    const int c = tm.v0 * 8;
    load8(((c < C0) ? x0 + c : x1 + (c - C0)) + ((size_t)b * HW + pf) * ((c < C0) ? C0 : C1), f0);
  }
  {
    // thread (stripe, cb): block column cb, slabs stripe, stripe + nstripe, ...  Consecutive threads read consecutive
    // float2 entries (coalesced); every sum runs in a fixed order (deterministic).
    const int nb0 = C0 / blk, nb1 = C1 / blk, nb = nb0 + nb1, bpg = cpg / blk;
    const int ncol = min(nb, 256);
    const int nstripe = max(1, min(8, 256 / ncol));
    const int stripe = threadIdx.x / ncol;
    if (stripe < nstripe) {
      const int cb = threadIdx.x % ncol;                          // nb <= 256 (checked by the launcher)
      const float2* src = (cb < nb0) ? reinterpret_cast<const float2*>(stats0) + (size_t)b * slots * nb0 + cb
                                     : reinterpret_cast<const float2*>(stats1) + (size_t)b * slots * nb1 + (cb - nb0);
      const int ld = (cb < nb0) ? nb0 : nb1;
      float a = 0.f, q = 0.f;
      // (all of a thread's slab loads in flight together: this fold is a latency chain at the head of every CTA)
#pragma unroll kGnFoldUnroll
      for (int sl = stripe; sl < slots; sl += nstripe) { const float2 v = __ldg(src + (size_t)sl * ld); a += v.x; q += v.y; }
      s_col[stripe][cb] = make_float2(a, q);
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
      const int g = threadIdx.x;
      float a = 0.f, q = 0.f;
      for (int i = 0; i < bpg; ++i) {
        const int cb = g * bpg + i;
        for (int st = 0; st < nstripe; ++st) { a += s_col[st][cb].x; q += s_col[st][cb].y; }
      }
      const float mean = a * inv_n;
      const float var = fmaxf(q * inv_n - mean * mean, 0.f);
      s_mean[g] = mean; s_rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  for (int v = tm.v0; v < nvec; v += tm.vstep) {
    const int c = v * 8;
    float sc[8], sf[8], g[8], be[8];
    if (v == tm.v0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { g[k] = g0[k]; be[k] = be0[k]; }
    } else {
      load8(gamma + c, g);
      load8(beta + c, be);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int grp = (c + k) / cpg;
      sc[k] = s_rstd[grp] * g[k];
      sf[k] = be[k] - s_mean[grp] * sc[k];
    }
    const __half* src = (c < C0) ? x0 + c : x1 + (c - C0);
    const int ld = (c < C0) ? C0 : C1;
#pragma unroll 4
    for (int p = p0 + tm.pofs; p < p1; p += tm.pstride) {
      const size_t pix = (size_t)b * HW + p;
      float f[8];
      if (v == tm.v0 && p == pf) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f0[k];
      } else {
        load8(src + pix * ld, f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float y = fmaf(f[k], sc[k], sf[k]);
        f[k] = do_silu ? __fdividef(y, 1.0f + __expf(-y)) : y;
      }
      store8(out + pix * C + c, f);
    }
  }
}

// Fold of the producers' block sums into per-(sample, group) totals {sum, sumsq} -- the input of gn_apply_kernel.  One CTA per
// sample (B x slots x nb float2 read ONCE, instead of once per apply CTA as gn_apply_blk_kernel does); fixed order.
__global__ void gn_fold_groups_kernel(int C0, int C1, int groups, const float* __restrict__ stats0, const float* __restrict__ stats1,
                                      int blk, int slots, float* __restrict__ out) {
  __shared__ float2 s_col[8][256];
  griddep_launch();
  const int C = C0 + C1, cpg = C / groups;
  const int b = blockIdx.x;
  const int nb0 = C0 / blk, nb1 = C1 / blk, nb = nb0 + nb1, bpg = cpg / blk;
  const int ncol = min(nb, 256);
  const int nstripe = max(1, min(8, 256 / ncol));
  const int stripe = threadIdx.x / ncol;
  griddep_wait();
  if (stripe < nstripe) {
    const int cb = threadIdx.x % ncol;
    const float2* src = (cb < nb0) ? reinterpret_cast<const float2*>(stats0) + (size_t)b * slots * nb0 + cb
                                   : reinterpret_cast<const float2*>(stats1) + (size_t)b * slots * nb1 + (cb - nb0);
    const int ld = (cb < nb0) ? nb0 : nb1;
    float a = 0.f, q = 0.f;
#pragma unroll 8
    for (int sl = stripe; sl < slots; sl += nstripe) { const float2 v = __ldg(src + (size_t)sl * ld); a += v.x; q += v.y; }
    s_col[stripe][cb] = make_float2(a, q);
  }
  __syncthreads();
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    float a = 0.f, q = 0.f;
    for (int i = 0; i < bpg; ++i) {
      const int cb = g * bpg + i;
      for (int st = 0; st < nstripe; ++st) { a += s_col[st][cb].x; q += s_col[st][cb].y; }
    }
    out[((size_t)b * groups + g) * 2] = a;
    out[((size_t)b * groups + g) * 2 + 1] = q;
  }
}

// Statistics fold for the GroupNorm that is applied INSIDE the consuming GEMM (gemm2_kernel<..., kXf = true>): the producers'
// block sums -> per (sample, channel) fp16 planes tab[b][0..2][C] = (mean_h, scale', shift') with
//   mean_h = fp16(mean_g),  scale' = k * rstd_g * gamma_c,  shift' = k * (beta_c - (mean_g - mean_h) * rstd_g * gamma_c),
// k = 1/2 when SiLU follows (silu(y) = y/2 * (1 + tanh(y/2))), else 1.  One CTA per sample; same fold order as
// gn_apply_blk_kernel (deterministic).  A few microseconds: B CTAs reading B * slots * nb float2.
__global__ void gn_fold_kernel(int C0, int C1, int HW, int groups, float eps, const float* __restrict__ stats0,
                               const float* __restrict__ stats1, int blk, int slots, const __half* __restrict__ gamma,
                               const __half* __restrict__ beta, int do_silu, __half* __restrict__ tab) {
  __shared__ float2 s_col[8][256];
  __shared__ float s_mean[64], s_rstd[64];
  griddep_launch();
  const int C = C0 + C1, cpg = C / groups;
  const int b = blockIdx.x;
  const float inv_n = 1.0f / ((float)cpg * (float)HW);
  griddep_wait();
  {
    const int nb0 = C0 / blk, nb1 = C1 / blk, nb = nb0 + nb1, bpg = cpg / blk;
    const int ncol = min(nb, 256);
    const int nstripe = max(1, min(8, 256 / ncol));
    const int stripe = threadIdx.x / ncol;
    if (stripe < nstripe) {
      const int cb = threadIdx.x % ncol;
      const float2* src = (cb < nb0) ? reinterpret_cast<const float2*>(stats0) + (size_t)b * slots * nb0 + cb
                                     : reinterpret_cast<const float2*>(stats1) + (size_t)b * slots * nb1 + (cb - nb0);
      const int ld = (cb < nb0) ? nb0 : nb1;
      float a = 0.f, q = 0.f;
#pragma unroll 8
      for (int sl = stripe; sl < slots; sl += nstripe) { const float2 v = __ldg(src + (size_t)sl * ld); a += v.x; q += v.y; }
      s_col[stripe][cb] = make_float2(a, q);
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
      const int g = threadIdx.x;
      float a = 0.f, q = 0.f;
      for (int i = 0; i < bpg; ++i) {
        const int cb = g * bpg + i;
        for (int st = 0; st < nstripe; ++st) { a += s_col[st][cb].x; q += s_col[st][cb].y; }
      }
      const float mean = a * inv_n;
      const float var = fmaxf(q * inv_n - mean * mean, 0.f);
      s_mean[g] = mean; s_rstd[g] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  const float k = do_silu ? 0.5f : 1.0f;
  __half* t0 = tab + (size_t)b * 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = s_mean[g], sc = s_rstd[g] * __half2float(__ldg(gamma + c));
    const __half mh = __float2half_rn(mean);
    t0[c] = mh;
    t0[C + c] = __float2half_rn(k * sc);
    t0[2 * C + c] = __float2half_rn(k * (__half2float(__ldg(beta + c)) - (mean - __half2float(mh)) * sc));
  }
}

// ------------------------------------------------------------------ LayerNorm fold (one-off, at weight finalisation)
// LN(x) W^T + b  ==  rstd * (x Wf^T - mean * colsum) + b32   with  Wf[n,c] = W[n,c]*gamma[c],
// colsum[n] = sum_c Wf[n,c] (of the ROUNDED fp16 Wf, so the mean term cancels exactly what the GEMM accumulates),
// b32[n] = b[n] + sum_c beta[c]*W[n,c].  One warp per weight row.
__global__ void fold_layernorm_kernel(const __half* __restrict__ W, const __half* __restrict__ bias,
                                      const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                      __half* __restrict__ Wf, float* __restrict__ colsum, float* __restrict__ b32,
                                      int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = __half2float(W[(size_t)n * K + k]);
    const __half wf = __float2half_rn(w * __half2float(gamma[k]));
    Wf[(size_t)n * K + k] = wf;
    cs += __half2float(wf);
    bs = fmaf(__half2float(beta[k]), w, bs);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { cs += __shfl_xor_sync(0xffffffffu, cs, o); bs += __shfl_xor_sync(0xffffffffu, bs, o); }
  if (lane == 0) { colsum[n] = cs; b32[n] = bs + (bias ? __half2float(bias[n]) : 0.f); }
}

// Per-row (sum, sumsq) of a [rows, C] fp16 matrix in the layout the GEMM epilogue produces ([rows][parts][2], part 0
// carries everything): used by the operator-level tests when no producing GEMM exists.
__global__ void row_stats_kernel(const __half* __restrict__ x, float* __restrict__ stats, int rows, int C, int parts) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f, ss = 0.f;
  for (int k = lane; k < C; k += 32) { const float v = __half2float(x[(size_t)row * C + k]); s += v; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if (lane == 0) {
    for (int i = 0; i < parts; ++i) { stats[((size_t)row * parts + i) * 2] = i ? 0.f : s; stats[((size_t)row * parts + i) * 2 + 1] = i ? 0.f : ss; }
  }
}

// ------------------------------------------------------------------ LayerNorm over the last dim (one warp per row)
template <int kMaxVec>
__global__ void layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, __half* __restrict__ out, int rows, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C / 8;
  float f[kMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int v = lane + j * 32;
    if (v < nvec) {
      load8(x + (size_t)row * C + v * 8, f[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += f[j][k];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int v = lane + j * 32;
    if (v < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = f[j][k] - mean; ss += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)C + eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j) {
    const int v = lane + j * 32;
    if (v < nvec) {
      float g[8], b[8], y[8];
      load8(gamma + v * 8, g);
      load8(beta + v * 8, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] = (f[j][k] - mean) * rstd * g[k] + b[k];
      store8(out + (size_t)row * C + v * 8, y);
    }
  }
}

// ------------------------------------------------------------------ timestep embedding + GEMV
// diffusers get_timestep_embedding(flip_sin_to_cos=True): [cos | sin], fp32 math, rounded to fp16 like `.to(dtype)`.
__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, __half* __restrict__ out, int B, int dim,
                                         float freq_shift, int flip) {
  const int half_dim = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half_dim) return;
  const int b = i / half_dim, k = i % half_dim;
  const float freq = expf(-logf(10000.0f) * (float)k / ((float)half_dim - freq_shift));
  const float a = t[b] * freq;
  const float s = sinf(a), c = cosf(a);
  __half* o = out + (size_t)b * dim;
  if (flip) { o[k] = __float2half_rn(c); o[half_dim + k] = __float2half_rn(s); }
  else      { o[k] = __float2half_rn(s); o[half_dim + k] = __float2half_rn(c); }
}

// out[b, n] = act_out( sum_k act_in(x[b,k]) * W[n,k] + bias[n] ), B <= 8.  A warp owns kRowsPerWarp consecutive output
// rows; the weight matrix is streamed exactly once with 16-byte loads (HBM-bound: the time-embedding MLP and the 22
// time_emb_proj layers); the activated input vector is staged once per block in shared memory as fp32.
constexpr int kGemvRowsPerWarp = 8;
__global__ void gemv_small_batch_kernel(const __half* __restrict__ x, int ldx, const __half* __restrict__ W,
                                        const __half* __restrict__ bias, __half* __restrict__ out, int ldo, int B,
                                        int N, int K, int silu_in, int silu_out) {
  extern __shared__ float xs[];  // [B][K]
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    float v = __half2float(x[(size_t)(i / K) * ldx + (i % K)]);
    xs[i] = silu_in ? silu_f(v) : v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n_base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kGemvRowsPerWarp;
#pragma unroll 2
  for (int r = 0; r < kGemvRowsPerWarp; ++r) {
    const int n = n_base + r;
    if (n >= N) break;
    float acc[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[b] = 0.f;
    for (int k = lane * 8; k < K; k += 32 * 8) {
      float w[8];
      load8(W + (size_t)n * K + k, w);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (b < B) {
          const float4 xa = *reinterpret_cast<const float4*>(xs + b * K + k);
          const float4 xb = *reinterpret_cast<const float4*>(xs + b * K + k + 4);
          acc[b] += w[0] * xa.x + w[1] * xa.y + w[2] * xa.z + w[3] * xa.w + w[4] * xb.x + w[5] * xb.y + w[6] * xb.z + w[7] * xb.w;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
#pragma unroll
      for (int o = 16; o; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
    }
    if (lane == 0) {
      for (int b = 0; b < B; ++b) {
        const float v = acc[b] + (bias ? __half2float(bias[n]) : 0.f);
        out[(size_t)b * ldo + n] = __float2half_rn(silu_out ? silu_f(v) : v);
      }
    }
  }
}

// ------------------------------------------------------------------ resampling / gathers
__global__ void upsample2x_nhwc_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W,
                                       int C) {
  const int nvec = C / 8;
  const size_t total = (size_t)B * (2 * H) * (2 * W) * nvec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = i % nvec; size_t t = i / nvec;
    const int x = t % (2 * W); t /= (2 * W);
    const int y = t % (2 * H); const int b = t / (2 * H);
    const Half8 val = *reinterpret_cast<const Half8*>(in + (((size_t)b * H + (y >> 1)) * W + (x >> 1)) * C + v * 8);
    *reinterpret_cast<Half8*>(out + i * 8) = val;
  }
}

// im2col for 3x3 / pad 1 / stride s on NHWC (C % 8 == 0): out[(b,yo,xo), tap*C + c].
__global__ void im2col3x3_nhwc_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W,
                                      int C, int stride, int Ho, int Wo) {
  const int nvec = C / 8;
  const size_t total = (size_t)B * Ho * Wo * 9 * nvec;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = i % nvec; size_t t = i / nvec;
    const int tap = t % 9; t /= 9;
    const int xo = t % Wo; t /= Wo;
    const int yo = t % Ho; const int b = t / Ho;
    const int y = yo * stride + tap / 3 - 1, x = xo * stride + tap % 3 - 1;
    Half8 val;
    if (y >= 0 && y < H && x >= 0 && x < W)
      val = *reinterpret_cast<const Half8*>(in + (((size_t)b * H + y) * W + x) * C + v * 8);
    else {
#pragma unroll
      for (int k = 0; k < 4; ++k) val.h[k] = __floats2half2_rn(0.f, 0.f);
    }
    *reinterpret_cast<Half8*>(out + i * 8) = val;
  }
}

// conv_in gather: NCHW fp16 sample [B,C,H,W] (C small, e.g. 4) -> [B*H*W, Kpad] with k = tap*C + c, zero padded.
__global__ void im2col_conv_in_kernel(const __half* __restrict__ in, __half* __restrict__ out, int B, int C, int H, int W,
                                      int Kpad) {
  const size_t total = (size_t)B * H * W * Kpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % Kpad; size_t t = i / Kpad;
    const int x = t % W; t /= W;
    const int y = t % H; const int b = t / H;
    __half val = __float2half_rn(0.f);
    if (k < 9 * C) {
      const int tap = k / C, c = k % C;
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = in[(((size_t)b * C + c) * H + yy) * W + xx];
    }
    out[i] = val;
  }
}

// ------------------------------------------------------------------ CFG combine + DDIM step (eta = 0)
// noise: [2*n_img, E] (rows [0,n_img) uncond, [n_img, 2 n_img) cond) or [n_img, E] when guidance is off.
// coef = {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)} read from device memory at index *step_idx so the
// kernel is CUDA-graph replayable.  prediction_type: 0 = epsilon, 1 = v_prediction.
__global__ void cfg_ddim_step_kernel(const __half* __restrict__ noise, __half* __restrict__ latents, size_t n_vec8,
                                     size_t half_offset_elems, float guidance, int use_cfg, int prediction_type,
                                     const float4* __restrict__ coef_table, const int* __restrict__ step_idx) {
  const float4 cf = coef_table[step_idx ? *step_idx : 0];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec8; i += (size_t)gridDim.x * blockDim.x) {
    float u[8], c[8], x[8];
    load8(noise + i * 8, u);
    load8(latents + i * 8, x);
    if (use_cfg) load8(noise + half_offset_elems + i * 8, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m = use_cfg ? (u[k] + guidance * (c[k] - u[k])) : u[k];
      float x0, eps;
      if (prediction_type == 0) { eps = m; x0 = (x[k] - cf.y * m) / cf.x; }
      else { x0 = cf.x * x[k] - cf.y * m; eps = cf.x * m + cf.y * x[k]; }
      x[k] = cf.z * x0 + cf.w * eps;
    }
    store8(latents + i * 8, x);
  }
}

// ------------------------------------------------------------------ weight packing (one-off, at load time)
// OIHW [O, I, 3, 3] -> [O, 9*Ipad] with k = tap*Ipad + c  (Ipad >= I, zero filled).
__global__ void pack_conv3x3_kernel(const __half* __restrict__ w, __half* __restrict__ out, int O, int I, int Ipad) {
  const size_t total = (size_t)O * 9 * Ipad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % Ipad; size_t t = i / Ipad;
    const int tap = t % 9; const int o = t / 9;
    out[i] = (c < I) ? w[((size_t)o * I + c) * 9 + tap] : __float2half_rn(0.f);
  }
}
// conv_in: OIHW [O, I, 3, 3] -> [O, Kpad] with k = tap*I + c.
// ResnetBlock2D's conv_shortcut (1x1) accumulated inside conv2 (3x3): one weight matrix [N][9*C + Cs] = [conv2 | shortcut] per
// row and one bias b2 + bs, so that out = conv2(h) + conv_shortcut(x) + biases is a single K loop over two tensors.
__global__ void concat_weight_rows_kernel(const __half* __restrict__ w0, int k0, const __half* __restrict__ w1, int k1,
                                          __half* __restrict__ out, int N) {
  const size_t kt = (size_t)k0 + k1, total = (size_t)N * kt;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / kt, k = i % kt;
    out[i] = k < (size_t)k0 ? w0[n * k0 + k] : w1[n * k1 + (k - k0)];
  }
}
__global__ void add_bias_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(__half2float(a[i]) + __half2float(b[i]));
}

// Nearest-x2 upsampling followed by a 3x3 convolution (diffusers Upsample2D) as four 2x2 convolutions on the LOW-resolution
// tensor, one per output phase (py, px) = parity of the output pixel: the 3 kernel rows that touch output row 2y + py read only
// two distinct input rows (py = 0: {y-1: ky 0; y: ky 1, 2}; py = 1: {y: ky 0, 1; y+1: ky 2}), likewise along x, so their
// weights can be summed beforehand -- 16 C N instead of 36 C N multiply-adds per low-resolution pixel, and the upsampled
// tensor is never written.  out[ph][n][(ty*2 + tx) * I + c] = sum_{ky in S(py,ty)} sum_{kx in S(px,tx)} w[n][c][ky][kx]
// (fp32 sum, rounded once to fp16), ph = py*2 + px.
__global__ void pack_upconv_kernel(const __half* __restrict__ w, __half* __restrict__ out, int O, int I) {
  const size_t total = (size_t)4 * O * 4 * I;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % I);
    size_t r = i / I;
    const int tap = (int)(r % 4); r /= 4;
    const int n = (int)(r % O);
    const int ph = (int)(r / O);
    const int py = ph >> 1, px = ph & 1, ty = tap >> 1, tx = tap & 1;
    // S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1}, S(1,1) = {2}
    const int ky0 = (py == 0) ? (ty == 0 ? 0 : 1) : (ty == 0 ? 0 : 2), ky1 = (py == 0) ? (ty == 0 ? 0 : 2) : (ty == 0 ? 1 : 2);
    const int kx0 = (px == 0) ? (tx == 0 ? 0 : 1) : (tx == 0 ? 0 : 2), kx1 = (px == 0) ? (tx == 0 ? 0 : 2) : (tx == 0 ? 1 : 2);
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc += __half2float(w[(((size_t)n * I + c) * 3 + ky) * 3 + kx]);
    out[i] = __float2half_rn(acc);
  }
}

__global__ void pack_conv_in_kernel(const __half* __restrict__ w, __half* __restrict__ out, int O, int I, int Kpad) {
  const size_t total = (size_t)O * Kpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % Kpad; const int o = i / Kpad;
    __half v = __float2half_rn(0.f);
    if (k < 9 * I) { const int tap = k / I, c = k % I; v = w[((size_t)o * I + c) * 9 + tap]; }
    out[i] = v;
  }
}
// GEGLU interleave: proj [2*inner, K] (rows [0,inner) value, [inner,2 inner) gate) -> tiles of `tile` rows:
// [value(hf rows) | gate(hf rows) | zero rows up to `tile`], zero padded to n_tiles*tile rows.  K == 1 packs a bias vector.
__global__ void pack_geglu_kernel(const __half* __restrict__ w, __half* __restrict__ out, int inner, int K, int tile,
                                  int hf, int n_tiles) {
  const size_t total = (size_t)n_tiles * tile * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % K; const int r = i / K;
    const int t = r / tile, rr = r % tile;
    const int j = t * hf + (rr % hf);
    const bool gate = rr >= hf;
    out[i] = (rr < 2 * hf && j < inner) ? w[((size_t)(gate ? inner + j : j)) * K + k] : __float2half_rn(0.f);
  }
}

// ------------------------------------------------------------------ CLIP text encoder helpers
// CLIPTextEmbeddings: out[b*S + t, :] = token_embedding[ids[b*S + t], :] + position_embedding[t, :].  8 channels / thread.
__global__ void clip_embed_kernel(const int* __restrict__ ids, const __half* __restrict__ tok, const __half* __restrict__ pos,
                                  __half* __restrict__ out, int rows, int S, int C, int vocab) {
  const int nvec = C / 8;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * nvec) return;
  const int r = (int)(i / nvec), v = (int)(i % nvec);
  int id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  float a[8], b[8];
  load8(tok + (size_t)id * C + v * 8, a);
  load8(pos + (size_t)(r % S) * C + v * 8, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] += b[k];
  store8(out + (size_t)r * C + v * 8, a);
}
// CLIP's `quick_gelu`: x * sigmoid(1.702 x), in place.
// act: 0 = quick_gelu x*sigmoid(1.702x) (OpenAI CLIP, SD-1.x text encoder), 1 = erf GELU (OpenCLIP ViT-H, SD-2.x text encoder)
__global__ void quick_gelu_kernel(__half* __restrict__ x, size_t nvec, int act) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float f[8];
    load8(x + i * 8, f);
    if (act == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = __fdividef(f[k], 1.0f + __expf(-1.702f * f[k]));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.5f * f[k] * (1.0f + erff(f[k] * 0.70710678118654752f));
    }
    store8(x + i * 8, f);
  }
}

// ---- CLIP image tower helpers
// Patch embedding as a GEMM: col[b*P + p, (c, ky, kx)] = image[b, c, py*ps + ky, px*ps + kx], K zero-padded to `kpad`.
__global__ void clip_patch_im2col_kernel(const __half* __restrict__ img, __half* __restrict__ col, int B, int Cin, int size, int ps,
                                         int kpad) {
  const int n = size / ps, P = n * n, kreal = Cin * ps * ps;
  const size_t total = (size_t)B * P * kpad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % kpad); const size_t rp = i / kpad;
    const int p = (int)(rp % P), b = (int)(rp / P);
    __half v = __float2half_rn(0.f);
    if (k < kreal) {
      const int c = k / (ps * ps), ky = (k / ps) % ps, kx = k % ps;
      const int y = (p / n) * ps + ky, x = (p % n) * ps + kx;
      v = img[(((size_t)b * Cin + c) * size + y) * size + x];
    }
    col[i] = v;
  }
}
// CLIPVisionEmbeddings: x[b, 0] = class_embedding + pos[0];  x[b, 1 + p] = patch[b, p] + pos[1 + p].
__global__ void clip_vision_embed_kernel(const __half* __restrict__ patch, const __half* __restrict__ cls, const __half* __restrict__ pos,
                                         __half* __restrict__ out, int B, int S, int C) {
  const int nvec = C / 8;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)B * S * nvec) return;
  const int v = (int)(i % nvec); const size_t r = i / nvec;
  const int t = (int)(r % S), b = (int)(r / S);
  float a[8], pe[8];
  if (t == 0) load8(cls + v * 8, a); else load8(patch + ((size_t)b * (S - 1) + (t - 1)) * C + v * 8, a);
  load8(pos + (size_t)t * C + v * 8, pe);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] += pe[k];
  store8(out + r * C + v * 8, a);
}
// out[i, :] = x[i * stride + idx[i], :]  (class token: idx == nullptr -> row i * stride; text pooling: idx = EOS position)
__global__ void gather_rows_kernel(const __half* __restrict__ x, const int* __restrict__ idx, __half* __restrict__ out, int n, int stride,
                                   int C) {
  const int nvec = C / 8;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)n * nvec) return;
  const int v = (int)(i % nvec), r = (int)(i / nvec);
  const size_t src = (size_t)r * stride + (idx ? idx[r] : 0);
  *reinterpret_cast<uint4*>(out + (size_t)r * C + v * 8) = *reinterpret_cast<const uint4*>(x + src * C + v * 8);
}
// logits_per_text[t, b] = exp(logit_scale) * <text_t / |text_t|, image_b / |image_b|>   (fp32; one warp per (t, b))
__global__ void clip_logits_kernel(const __half* __restrict__ txt, const __half* __restrict__ img, float* __restrict__ out, int T, int B,
                                   int D, float scale) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= T * B) return;
  const int t = w / B, b = w % B;
  float dot = 0.f, nt = 0.f, ni = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float a = __half2float(txt[(size_t)t * D + k]), c = __half2float(img[(size_t)b * D + k]);
    dot = fmaf(a, c, dot); nt = fmaf(a, a, nt); ni = fmaf(c, c, ni);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o); nt += __shfl_xor_sync(0xffffffffu, nt, o); ni += __shfl_xor_sync(0xffffffffu, ni, o);
  }
  if (lane == 0) out[w] = scale * dot * rsqrtf(nt) * rsqrtf(ni);
}

// ---- clip `preprocess` on the device (rows f3 / f4): PIL's antialiased bicubic resize in its own 8-bit fixed point
// (Pillow src/libImaging/Resample.c, ImagingResampleHorizontal_8bpc / Vertical_8bpc: coefficients quantised to 22 bits by the
// host, accumulator seeded with 1 << 21, result = clip8(acc >> 22), one pass per axis with a uint8 intermediate image), so
// the pixels CLIP sees are the ones `Image.open(png).resize(...)` would produce.
// One thread per output byte.  axis 0: along x (in [B,H,Win,C] -> out [B,H,Wout,C]); axis 1: along y.
__global__ void resample_u8_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int B, int Hin, int Win,
                                   int C, int Hout, int Wout, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                   int axis) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Hout * Wout * C;
  if (i >= total) return;
  const int c = (int)(i % C); size_t t = i / C;
  const int x = (int)(t % Wout); t /= Wout;
  const int y = (int)(t % Hout); const int b = (int)(t / Hout);
  const int o = axis == 0 ? x : y;
  const int lo = bounds[2 * o], n = bounds[2 * o + 1];
  const int* k = kk + (size_t)o * ksize;
  int acc = 1 << 21;
  if (axis == 0) {
    const unsigned char* src = in + (((size_t)b * Hin + y) * Win + lo) * C + c;
    for (int j = 0; j < n; ++j) acc += (int)src[(size_t)j * C] * k[j];
  } else {
    const unsigned char* src = in + (((size_t)b * Hin + lo) * Win + x) * C + c;
    for (int j = 0; j < n; ++j) acc += (int)src[(size_t)j * Win * C] * k[j];
  }
  const int v = acc >> 22;
  out[i] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// CenterCrop + ToTensor + Normalize: u8 [B,H,W,3] -> fp16 [B,3,n,n], ((v / 255) - mean) / std in fp32.
__global__ void clip_normalize_kernel(const unsigned char* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int top, int left,
                                      int n, float m0, float m1, float m2, float s0, float s1, float s2) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)B * 3 * n * n) return;
  const int x = (int)(i % n); size_t t = i / n;
  const int y = (int)(t % n); t /= n;
  const int c = (int)(t % 3); const int b = (int)(t / 3);
  const float v = (float)in[(((size_t)b * H + top + y) * W + left + x) * 3 + c] / 255.0f;
  const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  out[i] = __float2half_rn((v - m) / sd);
}

// ---- filteration mask compositing (DiverGen/filteration/get_clip_score.py:133-146, --use_mask):
//   mask_im = mask > 128;  image = image * mask_im + ones_like(image) * (1 - mask_im)   (background becomes the value 1, as there)
//   area    = sum(mask_im) / H / W   -- accumulated as an exact integer count per image (warp-aggregated atomics)
// img / out [B, HW, 3] uint8, mask [B, HW] uint8, count [B] (zeroed by the caller).
__global__ void mask_composite_u8_kernel(const unsigned char* __restrict__ img, const unsigned char* __restrict__ mask,
                                         unsigned char* __restrict__ out, unsigned int* __restrict__ count, int B, int HW) {
  const int b = blockIdx.y;
  unsigned int local = 0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const size_t i = (size_t)b * HW + p;
    const bool m = mask[i] > 128;
    local += m ? 1u : 0u;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = m ? img[i * 3 + c] : (unsigned char)1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count + b, local);
}

// ------------------------------------------------------------------ VAE decoder helpers
// post_quant_conv: 1x1 conv over the latent channels (C <= 8) on NCHW fp16, with the 1/scaling_factor of
// `vae.decode(latents / scaling_factor)` folded in.  out[b,co,p] = bias[co] + sum_ci W[co,ci] * (z[b,ci,p] * scale).
__global__ void latent_pointwise_kernel(const __half* __restrict__ z, const __half* __restrict__ W, const __half* __restrict__ bias,
                                        __half* __restrict__ out, int B, int C, int HW, float scale) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  const int b = i / HW, p = i % HW;
  float v[8];
  for (int c = 0; c < C; ++c) v[c] = __half2float(z[((size_t)b * C + c) * HW + p]) * scale;
  for (int co = 0; co < C; ++co) {
    float a = __half2float(bias[co]);
    for (int ci = 0; ci < C; ++ci) a = fmaf(__half2float(W[co * C + ci]), v[ci], a);
    out[((size_t)b * C + co) * HW + p] = __float2half_rn(a);
  }
}

// Row softmax of an fp16 score matrix in place: P[r, :] = softmax(S[r, :] * scale), fp32 math, one CTA per row
// (the single-head, full-width attention of the VAE mid block; rows of 4096 scores at 512 x 512 output).
__global__ void softmax_rows_kernel(__half* __restrict__ S, int n, float scale_log2) {
  __shared__ float red[32];
  __half* row = S + (size_t)blockIdx.x * n;
  float mx = -INFINITY;
  for (int i = threadIdx.x * 8; i < n; i += blockDim.x * 8) {
    float f[8];
    load8(row + i, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) mx = fmaxf(mx, f[k]);
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float m = mx * scale_log2;
  float sum = 0.f;
  for (int i = threadIdx.x * 8; i < n; i += blockDim.x * 8) {
    float f[8];
    load8(row + i, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += exp2f(fmaf(f[k], scale_log2, -m));
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[w];
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x * 8; i < n; i += blockDim.x * 8) {
    float f[8];
    load8(row + i, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = exp2f(fmaf(f[k], scale_log2, -m)) * inv;
    store8(row + i, f);
  }
}

// [rows, C] -> [C, rows] fp16 (V^T as the K-major "weight" operand of the P*V GEMM), 32x32 tiles through shared memory.
__global__ void transpose_rows_kernel(const __half* __restrict__ in, __half* __restrict__ out, int rows, int C) {
  __shared__ __half tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < C) ? in[(size_t)r * C + c] : __float2half_rn(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

}  // namespace dg
