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
