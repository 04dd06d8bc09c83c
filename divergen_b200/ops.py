"""Single-operator entry points of the C ABI on torch CUDA tensors (used by the parity tests and the benches).

Each function is a thin ctypes call of the same launcher the UNet uses; tensors are fp16, contiguous, NHWC / row-major.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _env(t: torch.Tensor):
    if t.device.type != "cuda":
        raise ValueError("divergen_b200 has no CPU path")
    return _lib.load(), _lib.context(t.device.index or 0), C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk16(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.float16 or not t.is_contiguous()):
            raise ValueError("tensors must be contiguous fp16")


def linear(x, weight, bias=None, residual=None):
    """out[M, N] = x[M, K] @ weight[N, K]^T (+ bias) (+ residual)."""
    _chk16(x, weight, bias, residual)
    lib, ctx, s = _env(x)
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_gemm(ctx, _p(x), _p(weight), _p(bias), _p(residual), _p(out), M, K, N, N, 0, s), "dg_op_gemm")
    return out


def geglu_linear(x, proj_weight, proj_bias):
    """GEGLU: h, g = (x @ W^T + b).chunk(2); out = h * gelu(g).  W: [2*inner, K]."""
    _chk16(x, proj_weight, proj_bias)
    lib, ctx, s = _env(x)
    M, K = x.shape
    inner = proj_weight.shape[0] // 2
    rows = lib.dg_op_geglu_packed_rows(inner)
    wp = torch.empty((rows, K), dtype=torch.float16, device=x.device)
    bp = torch.empty((rows,), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_pack_geglu(ctx, _p(proj_weight), _p(proj_bias), _p(wp), _p(bp), inner, K, s), "dg_op_pack_geglu")
    out = torch.empty((M, inner), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_gemm(ctx, _p(x), _p(wp), _p(bp), _p(None), _p(out), M, K, rows, inner, 1, s), "dg_op_gemm(geglu)")
    return out


def conv3x3_nhwc(x, weight_oihw, bias=None, x1=None, rowvec=None, residual=None):
    """3x3 / stride 1 / pad 1 convolution on NHWC x [B,H,W,C0] (optionally channel-concatenated with x1)."""
    _chk16(x, weight_oihw, bias, x1, rowvec, residual)
    lib, ctx, s = _env(x)
    B, H, W, C0 = x.shape
    C1 = x1.shape[3] if x1 is not None else 0
    O, I = weight_oihw.shape[:2]
    assert I == C0 + C1
    wp = torch.empty((O, 9 * I), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_pack_conv3x3(ctx, _p(weight_oihw), _p(wp), O, I, s), "dg_op_pack_conv3x3")
    ldo = (O + 7) // 8 * 8   # TMA store needs a 16-byte row pitch
    if residual is not None and ldo != O:
        raise ValueError("residual needs an output width that is a multiple of 8")
    out = torch.zeros((B, H, W, ldo), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_conv3x3(ctx, _p(x), C0, _p(x1), C1, _p(wp), _p(bias), _p(rowvec),
                                 rowvec.stride(0) if rowvec is not None else 0, _p(residual), _p(out), B, H, W, O, ldo, s),
               "dg_op_conv3x3")
    return out[..., :O]


def attention(q, k, v, heads: int):
    """q [B,Sq,heads*d], k/v [B,Sk,heads*d] (may be strided column views of a fused QKV matrix)."""
    for t in (q, k, v):
        if t.dtype != torch.float16 or t.stride(2) != 1:
            raise ValueError("q/k/v must be fp16 with unit inner stride")
    lib, ctx, s = _env(q)
    B, Sq, Cq = q.shape
    Sk = k.shape[1]
    d = Cq // heads
    for t, S in ((q, Sq), (k, Sk), (v, Sk)):
        if t.stride(0) != S * t.stride(1):
            raise ValueError("batch stride must equal S * row stride")
    out = torch.empty((B, Sq, Cq), dtype=torch.float16, device=q.device)
    _lib.check(lib.dg_op_attention(ctx, _p(q), q.stride(1), _p(k), k.stride(1), _p(v), v.stride(1), _p(out), B, heads,
                                   Sq, Sk, d, s), "dg_op_attention")
    return out


def groupnorm_nhwc(x, gamma, beta, groups: int, eps: float, silu: bool, x1=None):
    _chk16(x, gamma, beta, x1)
    lib, ctx, s = _env(x)
    B, H, W, C0 = x.shape
    C1 = x1.shape[3] if x1 is not None else 0
    out = torch.empty((B, H, W, C0 + C1), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_groupnorm(ctx, _p(x), C0, _p(x1), C1, _p(gamma), _p(beta), _p(out), B, H * W, groups, eps,
                                   int(silu), s), "dg_op_groupnorm")
    return out


def layernorm(x, gamma, beta, eps: float = 1e-5):
    _chk16(x, gamma, beta)
    lib, ctx, s = _env(x)
    rows, Cc = x.shape
    out = torch.empty_like(x)
    _lib.check(lib.dg_op_layernorm(ctx, _p(x), _p(gamma), _p(beta), _p(out), rows, Cc, eps, s), "dg_op_layernorm")
    return out


def time_embedding(timesteps, dim: int, w1, b1, w2, b2):
    _chk16(w1, b1, w2, b2)
    lib, ctx, s = _env(w1)
    B = len(timesteps)
    temb = w1.shape[0]
    out = torch.empty((B, temb), dtype=torch.float16, device=w1.device)
    arr = (C.c_float * B)(*[float(t) for t in timesteps])
    _lib.check(lib.dg_op_time_embedding(ctx, arr, B, dim, temb, _p(w1), _p(b1), _p(w2), _p(b2), _p(out), s),
               "dg_op_time_embedding")
    return out


def cfg_ddim_step(noise_pred, latents, alpha_t: float, alpha_prev: float, guidance_scale: float, prediction_type: str):
    """In place on `latents`: CFG combine (when guidance_scale > 1, noise_pred holds [uncond; cond]) + DDIM step."""
    _chk16(noise_pred, latents)
    lib, ctx, s = _env(latents)
    n = latents.shape[0]
    _lib.check(lib.dg_cfg_ddim_step(ctx, _p(noise_pred), _p(latents), n, latents.numel() // n, alpha_t, alpha_prev,
                                    guidance_scale, {"epsilon": 0, "v_prediction": 1}[prediction_type], s),
               "dg_cfg_ddim_step")
    return latents


# ------------------------------------------------------------------ fused-epilogue forms (what the UNet launches)
def _pf(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def row_parts(n_out: int) -> int:
    return int(_lib.load().dg_op_gemm_row_parts(n_out))


def linear_stats(x, weight, bias=None, residual=None, gn_blk: int = 0, hw: int = 0):
    """linear() that also returns the fused statistics of its fp16 result: per-row (sum, sumsq) [M, 2] (partials already
    added up) and, when gn_blk > 0, GroupNorm sums per 32-row slab and gn_blk-channel block [M // hw, hw // 32, N // gn_blk, 2]."""
    _chk16(x, weight, bias, residual)
    lib, ctx, s = _env(x)
    M, K = x.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.float16, device=x.device)
    parts = row_parts(N)
    rs = torch.full((M, parts, 2), float("nan"), dtype=torch.float32, device=x.device)
    gs = torch.full((M // hw, hw // 32, N // gn_blk, 2), float("nan"), dtype=torch.float32, device=x.device) if gn_blk else None
    _lib.check(lib.dg_op_gemm_fused(ctx, _p(x), _p(weight), _p(bias), _pf(None), _pf(None), _pf(None), 0, 0.0,
                                    _p(residual), _p(out), M, K, N, N, 0, _pf(rs), _pf(gs), gn_blk, hw, s),
               "dg_op_gemm_fused(stats)")
    return out, rs.sum(1), gs


def layernorm_linear(x, gamma, beta, weight, bias=None, eps: float = 1e-5, geglu: bool = False, row_stats=None):
    """LayerNorm(x) @ weight^T + bias with the normalisation folded into the GEMM epilogue.  `row_stats` [M, parts, 2]
    may come from the GEMM that produced x (linear_stats' raw partials); otherwise dg_op_row_stats computes them."""
    _chk16(x, gamma, beta, weight, bias)
    lib, ctx, s = _env(x)
    M, K = x.shape
    if geglu:
        inner = weight.shape[0] // 2
        rows = lib.dg_op_geglu_packed_rows(inner)
        wp = torch.empty((rows, K), dtype=torch.float16, device=x.device)
        bp = torch.empty((rows,), dtype=torch.float16, device=x.device)
        _lib.check(lib.dg_op_pack_geglu(ctx, _p(weight), _p(bias), _p(wp), _p(bp), inner, K, s), "dg_op_pack_geglu")
        weight, bias, n_w, n_out = wp, bp, rows, inner
    else:
        n_w = n_out = weight.shape[0]
    wf = torch.empty_like(weight)
    cs = torch.empty((n_w,), dtype=torch.float32, device=x.device)
    b32 = torch.empty((n_w,), dtype=torch.float32, device=x.device)
    _lib.check(lib.dg_op_fold_layernorm(ctx, _p(weight), _p(bias), _p(gamma), _p(beta), _p(wf), _pf(cs), _pf(b32), n_w, K, s),
               "dg_op_fold_layernorm")
    if row_stats is None:
        parts = row_parts(K)
        row_stats = torch.empty((M, parts, 2), dtype=torch.float32, device=x.device)
        _lib.check(lib.dg_op_row_stats(ctx, _p(x), _pf(row_stats), M, K, parts, s), "dg_op_row_stats")
    out = torch.empty((M, n_out), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_gemm_fused(ctx, _p(x), _p(wf), _p(None), _pf(b32), _pf(cs), _pf(row_stats), K, eps, _p(None), _p(out),
                                    M, K, n_w, n_out, int(geglu), _pf(None), _pf(None), 0, 0, s), "dg_op_gemm_fused(ln)")
    return out


def conv3x3_stats(x, weight_oihw, bias, gn_blk: int):
    """conv3x3_nhwc() that also returns the GroupNorm slab/block sums [B, H*W // 32, N // gn_blk, 2] of its fp16 result."""
    _chk16(x, weight_oihw, bias)
    lib, ctx, s = _env(x)
    B, H, W, C0 = x.shape
    O = weight_oihw.shape[0]
    wp = torch.empty((O, 9 * C0), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_pack_conv3x3(ctx, _p(weight_oihw), _p(wp), O, C0, s), "dg_op_pack_conv3x3")
    out = torch.empty((B, H, W, O), dtype=torch.float16, device=x.device)
    gs = torch.full((B, H * W // 32, O // gn_blk, 2), float("nan"), dtype=torch.float32, device=x.device)
    _lib.check(lib.dg_op_conv3x3_stats(ctx, _p(x), C0, _p(wp), _p(bias), _p(out), B, H, W, O, _pf(gs), gn_blk, s),
               "dg_op_conv3x3_stats")
    return out, gs


def groupnorm_fused_nhwc(x, stats, gamma, beta, groups: int, eps: float, silu: bool, blk: int, x1=None, stats1=None):
    """GroupNorm apply (+SiLU) from block sums produced by the GEMM epilogues (one stats array per source)."""
    _chk16(x, gamma, beta, x1)
    lib, ctx, s = _env(x)
    B, H, W, C0 = x.shape
    C1 = x1.shape[3] if x1 is not None else 0
    out = torch.empty((B, H, W, C0 + C1), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_groupnorm_fused(ctx, _p(x), C0, _pf(stats), _p(x1), C1, _pf(stats1), blk, _p(gamma), _p(beta), _p(out),
                                         B, H * W, groups, eps, int(silu), s), "dg_op_groupnorm_fused")
    return out


def conv_gn(x, stats, gamma, beta, groups: int, eps: float, silu: bool, blk: int, weight, bias=None, x1=None, stats1=None,
            residual=None):
    """conv(act(norm(x))) with the GroupNorm (+ SiLU) applied inside the GEMM's operand path (dg_op_conv3x3_gn).
    weight [O, I, 3, 3] -> 3x3 conv; weight [O, I] -> 1x1 conv / Linear over the pixels.  stats / stats1: block sums
    [B, H*W // 32, C // blk, 2] of x / x1 (any assignment of a sample's pixels to its 32-pixel slabs)."""
    _chk16(x, gamma, beta, weight, bias, x1, residual)
    lib, ctx, s = _env(x)
    B, H, W, C0 = x.shape
    C1 = x1.shape[3] if x1 is not None else 0
    O, I = weight.shape[:2]
    assert I == C0 + C1
    taps = 9 if weight.dim() == 4 else 1
    if taps == 9:
        wp = torch.empty((O, 9 * I), dtype=torch.float16, device=x.device)
        _lib.check(lib.dg_op_pack_conv3x3(ctx, _p(weight), _p(wp), O, I, s), "dg_op_pack_conv3x3")
    else:
        wp = weight.contiguous()
    ldo = (O + 7) // 8 * 8   # TMA store needs a 16-byte row pitch
    if residual is not None and ldo != O:
        raise ValueError("residual needs an output width that is a multiple of 8")
    out = torch.zeros((B, H, W, ldo), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_conv3x3_gn(ctx, _p(x), C0, _pf(stats), _p(x1), C1, _pf(stats1), blk, _p(gamma), _p(beta), groups, eps,
                                    int(silu), _p(wp), _p(bias), _p(residual), _p(out), B, H, W, O, ldo, taps, s), "dg_op_conv3x3_gn")
    return out[..., :O]


def conv3x3_shortcut(x, weight_oihw, bias, xs0, shortcut_weight, shortcut_bias, xs1=None):
    """conv3x3(x) + conv1x1(cat[xs0, xs1]) + biases in one K loop (dg_op_conv3x3_shortcut): ResnetBlock2D's conv2 with its
    conv_shortcut.  shortcut_weight [N, Cs0 + Cs1]."""
    _chk16(x, weight_oihw, bias, xs0, shortcut_weight, shortcut_bias, xs1)
    lib, ctx, s = _env(x)
    B, H, W, Cin = x.shape
    O = weight_oihw.shape[0]
    wp = torch.empty((O, 9 * Cin), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_pack_conv3x3(ctx, _p(weight_oihw), _p(wp), O, Cin, s), "dg_op_pack_conv3x3")
    wcat = torch.cat([wp, shortcut_weight.reshape(O, -1)], dim=1).contiguous()
    b = (bias.float() + shortcut_bias.float()).half()
    out = torch.empty((B, H, W, O), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_conv3x3_shortcut(ctx, _p(x), Cin, _p(wcat), _p(b), _p(xs0), xs0.shape[3], _p(xs1),
                                          xs1.shape[3] if xs1 is not None else 0, _p(out), B, H, W, O, s), "dg_op_conv3x3_shortcut")
    return out


def conv3x3_stride2(x, weight_oihw, bias):
    """Downsample2D: conv3x3 / stride 2 / pad 1 on NHWC x [B, H, W, C] (H, W even) -> [B, H/2, W/2, N] (dg_op_conv3x3_stride2)."""
    _chk16(x, weight_oihw, bias)
    lib, ctx, s = _env(x)
    B, H, W, Cin = x.shape
    O = weight_oihw.shape[0]
    wp = torch.empty((O, 9 * Cin), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_pack_conv3x3(ctx, _p(weight_oihw), _p(wp), O, Cin, s), "dg_op_pack_conv3x3")
    out = torch.empty((B, H // 2, W // 2, O), dtype=torch.float16, device=x.device)
    _lib.check(lib.dg_op_conv3x3_stride2(ctx, _p(x), Cin, _p(wp), _p(bias), _p(out), B, H // 2, W // 2, O, _pf(None), 0, s),
               "dg_op_conv3x3_stride2")
    return out


def upsample_conv3x3(x, weight_oihw, bias, gn_blk: int = 0):
    """Upsample2D (nearest x2 + conv3x3) on NHWC x [B, H, W, C] -> [B, 2H, 2W, N] through the four-phase decomposition
    (dg_op_upsample_conv3x3); with gn_blk also the GroupNorm block sums [B, 4*H*W // 32, N // gn_blk, 2] of the result."""
    _chk16(x, weight_oihw, bias)
    lib, ctx, s = _env(x)
    B, H, W, Cin = x.shape
    O = weight_oihw.shape[0]
    out = torch.empty((B, 2 * H, 2 * W, O), dtype=torch.float16, device=x.device)
    gs = torch.full((B, 4 * H * W // 32, O // gn_blk, 2), float("nan"), dtype=torch.float32, device=x.device) if gn_blk else None
    _lib.check(lib.dg_op_upsample_conv3x3(ctx, _p(x), Cin, _p(weight_oihw.contiguous()), _p(bias), _p(out), B, H, W, O, _pf(gs), gn_blk, s),
               "dg_op_upsample_conv3x3")
    return (out, gs) if gn_blk else out


def image_to_uint8(images):
    """[B, C, H, W] fp16 in [-1, 1] -> [B, H, W, C] uint8 on the device: diffusers.utils.pt_to_pil's arithmetic
    (`(x / 2 + 0.5).clamp(0, 1)`, `* 255`, round) without the host round trip of the float image."""
    _chk16(images)
    lib, ctx, s = _env(images)
    B, Cc, H, W = images.shape
    out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=images.device)
    _lib.check(lib.dg_op_image_to_uint8(ctx, _p(images), C.c_void_p(out.data_ptr()), B, Cc, H, W, s), "dg_op_image_to_uint8")
    return out
