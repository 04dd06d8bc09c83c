"""clip's `preprocess` for images that are already on the device: `Resize(n_px, BICUBIC)`,
`CenterCrop(n_px)`, `ToTensor`, `Normalize(mean, std)` as in DiverGen/filteration/get_clip_score.py:128-150 (OpenAI clip's
`_transform`), with the resize done the way Pillow does it so that the generated image can be scored before it is written.

Pillow (third-party, version in this image: see `PIL.__version__`; src/libImaging/Resample.c) resamples 8-bit images in two
passes (x, then y) with antialiased filter taps quantised to 22-bit fixed point; `resample_coeffs` restates its
`precompute_coeffs` + `normalize_coeffs_8bpc` in double precision on the host (pure index / weight arithmetic; `tests/test_preprocess.py` pins it against PIL
itself on the CPU), the two integer passes run in `resample_u8_kernel`.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=64)
def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's precompute_coeffs(in0=0, in1=in_size, BICUBIC) + normalize_coeffs_8bpc -> (bounds [out,2] int32,
    coeffs [out, ksize] int32, ksize)."""
    support = 2.0
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * xmax
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
        for x in range(xmax):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
    return bounds, kk, ksize


def resize_size(h: int, w: int, n_px: int) -> Tuple[int, int]:
    """torchvision `Resize(n_px)`: the shorter side becomes n_px, the longer int(n_px * long / short)."""
    if w <= h:
        return int(n_px * h / w), n_px
    return n_px, int(n_px * w / h)


def _dev_tables(in_size: int, out_size: int, device):
    bounds, kk, ksize = resample_coeffs(in_size, out_size)
    return torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ksize


def resize_u8(images: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """[B, H, W, C] uint8 (device) -> [B, out_h, out_w, C] uint8, bit-identical to `PIL.Image.resize(..., BICUBIC)`."""
    if images.dtype != torch.uint8 or images.dim() != 4 or not images.is_cuda:
        raise ValueError("images must be a [B, H, W, C] uint8 CUDA tensor")
    lib, ctx = _lib.load(), _lib.context(images.device.index or 0)
    stream = C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)
    x = images.contiguous()
    b, h, w, c = x.shape
    ip = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_int32))
    if out_w != w:
        bd, kk, ks = _dev_tables(w, out_w, x.device)
        y = torch.empty((b, h, out_w, c), dtype=torch.uint8, device=x.device)
        _lib.check(lib.dg_op_resample_u8(ctx, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), b, h, w, c, h, out_w, ip(bd), ip(kk),
                                         ks, 0, stream), "dg_op_resample_u8(x)")
        x, w = y, out_w
    if out_h != h:
        bd, kk, ks = _dev_tables(h, out_h, x.device)
        y = torch.empty((b, out_h, w, c), dtype=torch.uint8, device=x.device)
        _lib.check(lib.dg_op_resample_u8(ctx, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), b, h, w, c, out_h, w, ip(bd), ip(kk),
                                         ks, 1, stream), "dg_op_resample_u8(y)")
        x = y
    return x


def clip_preprocess(images: torch.Tensor, n_px: int = 224, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
    """[B, H, W, 3] uint8 (device, e.g. the pipeline's `output_type='uint8'`) -> [B, 3, n_px, n_px] fp16 `pixel_values`."""
    b, h, w, c = images.shape
    if c != 3:
        raise ValueError("clip_preprocess expects RGB images")
    oh, ow = resize_size(h, w, n_px)
    x = resize_u8(images, oh, ow)
    top, left = int(round((oh - n_px) / 2.0)), int(round((ow - n_px) / 2.0))
    out = torch.empty((b, 3, n_px, n_px), dtype=torch.float16, device=images.device)
    lib, ctx = _lib.load(), _lib.context(images.device.index or 0)
    m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _lib.check(lib.dg_op_clip_normalize(ctx, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), b, oh, ow, top, left, n_px, m, s,
                                        C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)), "dg_op_clip_normalize")
    return out


def mask_composite(images: torch.Tensor, masks: torch.Tensor):
    """filteration/get_clip_score.py:133-146 on the device: `mask_im = mask > 128`, background pixels become the value 1,
    `area = sum(mask_im) / H / W`.  images [B, H, W, 3] uint8, masks [B, H, W] uint8 -> (composited images, areas fp64 [B])."""
    if images.dtype != torch.uint8 or masks.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
        raise ValueError("mask_composite expects [B, H, W, 3] uint8 images and [B, H, W] uint8 masks")
    if tuple(masks.shape) != tuple(images.shape[:3]):
        raise ValueError("mask shape {} does not match images {}".format(tuple(masks.shape), tuple(images.shape)))
    images, masks = images.contiguous(), masks.contiguous()
    b, h, w, _ = images.shape
    out = torch.empty_like(images)
    count = torch.empty(b, dtype=torch.int32, device=images.device)
    lib, ctx = _lib.load(), _lib.context(images.device.index or 0)
    _lib.check(lib.dg_op_mask_composite_u8(ctx, C.c_void_p(images.data_ptr()), C.c_void_p(masks.data_ptr()), C.c_void_p(out.data_ptr()),
                                           C.c_void_p(count.data_ptr()), b, h, w,
                                           C.c_void_p(torch.cuda.current_stream(images.device).cuda_stream)), "dg_op_mask_composite_u8")
    return out, count.cpu().double() / h / w
