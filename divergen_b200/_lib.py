"""ctypes binding of libdivergen_b200.so (include/divergen_b200.h).

The library is the product: if it is missing, or the device is not sm_100, loading fails loudly.
There is no CPU / PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# DG_LIB_PATH: load another build of the same sources (A/B runs of compile-time switches); the default is the in-tree library
LIB_PATH = os.environ.get("DG_LIB_PATH") or os.path.join(_HERE, "libdivergen_b200.so")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/dg_api.cu -> libdivergen_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(os.path.dirname(_HERE), "include", "divergen_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("DG_NVCC_EXTRA", "").split()      # e.g. -DDG_GEMM_STAMPS for the clock-stamped debug build
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", LIB_PATH, os.path.join(CSRC, "dg_api.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


class DgError(RuntimeError):
    pass


class UNetConfigC(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("sample_size", C.c_int32),
                ("block_out_channels", C.c_int32 * 4), ("layers_per_block", C.c_int32), ("num_heads", C.c_int32 * 4),
                ("cross_attention_dim", C.c_int32), ("norm_num_groups", C.c_int32), ("norm_eps", C.c_float),
                ("use_linear_projection", C.c_int32), ("upcast_attention", C.c_int32), ("down_has_attn", C.c_int32 * 4),
                ("flip_sin_to_cos", C.c_int32), ("freq_shift", C.c_float)]


# every symbol include/divergen_b200.h declares: (name, restype, argtypes)
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "dg_version": (_I, []),
    "dg_plan_attention_grid": (_I, [_I, _I, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "dg_last_error": (C.c_char_p, []),
    "dg_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "dg_ctx_destroy": (None, [_P]),
    "dg_unet_create": (_I, [_P, C.POINTER(UNetConfigC), C.POINTER(_P)]),
    "dg_unet_destroy": (None, [_P]),
    "dg_unet_num_weights": (_I, [_P]),
    "dg_unet_weight_name": (C.c_char_p, [_P, _I]),
    "dg_unet_weight_shape": (_I, [_P, _I, C.POINTER(_L), C.POINTER(_I)]),
    "dg_unet_set_weight": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(_L)]),
    "dg_unet_missing_weights": (_I, [_P]),
    "dg_unet_prepare": (_I, [_P, _I, _I, _I, _I]),
    "dg_unet_forward": (_I, [_P, _P, C.POINTER(_F), _I, _P, _I, _P, _I, _I, _I, _P]),
    "dg_unet_set_graphs": (_I, [_P, _I]),
    "dg_unet_last_launch_count": (_L, [_P]),
    "dg_unet_set_family_mask": (_I, [_P, _I]),
    "dg_unet_profile_forward": (_I, [_P, _P, C.POINTER(_F), _I, _P, _I, _P, _I, _I, _I, _P, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_L), C.POINTER(C.c_double)]),
    "dg_cfg_ddim_step": (_I, [_P, _P, _P, _I, _L, _F, _F, _F, _I, _P]),
    "dg_denoise_loop": (_I, [_P, _P, _P, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), C.POINTER(_F), _I, _F, _I, _P]),
    "dg_vae_create": (_I, [_P, C.POINTER(_I), _I, C.POINTER(_P)]),
    "dg_vae_destroy": (None, [_P]),
    "dg_vae_num_weights": (_I, [_P]),
    "dg_vae_weight_name": (C.c_char_p, [_P, _I]),
    "dg_vae_weight_shape": (_I, [_P, _I, C.POINTER(_L), C.POINTER(_I)]),
    "dg_vae_set_weight": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(_L)]),
    "dg_vae_prepare": (_I, [_P, _I, _I, _I]),
    "dg_vae_decode": (_I, [_P, _P, _F, _P, _I, _I, _I, _P]),
    "dg_clip_create": (_I, [_P, _I, _I, _I, _I, _I, _I, C.POINTER(_P)]),
    "dg_clip_destroy": (None, [_P]),
    "dg_clip_num_weights": (_I, [_P]),
    "dg_clip_weight_name": (C.c_char_p, [_P, _I]),
    "dg_clip_weight_shape": (_I, [_P, _I, C.POINTER(_L), C.POINTER(_I)]),
    "dg_clip_set_weight": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(_L)]),
    "dg_clip_set_activation": (_I, [_P, _I]),
    "dg_clip_prepare": (_I, [_P, _I]),
    "dg_clip_encode": (_I, [_P, C.POINTER(_I), _I, _I, _P, _P]),
    "dg_clipscore_create": (_I, [_P, C.POINTER(_I), C.POINTER(_I), _I, C.POINTER(_P)]),
    "dg_clipscore_destroy": (None, [_P]),
    "dg_clipscore_num_weights": (_I, [_P]),
    "dg_clipscore_weight_name": (C.c_char_p, [_P, _I]),
    "dg_clipscore_weight_shape": (_I, [_P, _I, C.POINTER(_L), C.POINTER(_I)]),
    "dg_clipscore_set_weight": (_I, [_P, C.c_char_p, _P, _I, C.POINTER(_L)]),
    "dg_clipscore_set_logit_scale": (_I, [_P, _F]),
    "dg_clipscore_prepare": (_I, [_P, _I, _I]),
    "dg_clipscore_score": (_I, [_P, _P, _I, C.POINTER(_I), C.POINTER(_I), _I, _I, C.POINTER(_F), _P]),
    "dg_op_image_to_uint8": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "dg_op_mask_composite_u8": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "dg_op_resample_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), _I, _I, _P]),
    "dg_op_clip_normalize": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _P]),
    "dg_op_gemm": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "dg_op_pack_geglu": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "dg_op_geglu_packed_rows": (_I, [_I]),
    "dg_op_gemm_row_parts": (_I, [_I]),
    "dg_op_gemm_fused": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _F, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "dg_op_fold_layernorm": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "dg_op_row_stats": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "dg_op_conv3x3_stats": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "dg_op_groupnorm_fused": (_I, [_P, _P, _I, _P, _P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _F, _I, _P]),
    "dg_op_pack_conv3x3": (_I, [_P, _P, _P, _I, _I, _P]),
    "dg_op_conv3x3": (_I, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "dg_op_conv3x3_gn": (_I, [_P, _P, _I, _P, _P, _I, _P, _I, _P, _P, _I, _F, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "dg_op_conv3x3_shortcut": (_I, [_P, _P, _I, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "dg_op_conv3x3_stride2": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "dg_op_upsample_conv3x3": (_I, [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "dg_op_attention": (_I, [_P, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "dg_op_groupnorm": (_I, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _F, _I, _P]),
    "dg_op_layernorm": (_I, [_P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "dg_op_time_embedding": (_I, [_P, C.POINTER(_F), _I, _I, _I, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (never builds implicitly on a GPU box: the .so ships with the tree)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DgError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dg_last_error().decode("utf-8", "replace")
        exc = ValueError if rc in (-1, -2, -3) else RuntimeError
        raise exc(f"divergen_b200 {what} failed (code {rc}): {msg}")


_ctx_cache = {}


def context(device_index: int):
    """One dg_ctx per device per process."""
    if device_index not in _ctx_cache:
        h = _P()
        check(load().dg_ctx_create(device_index, C.byref(h)), "dg_ctx_create")
        _ctx_cache[device_index] = h
    return _ctx_cache[device_index]
