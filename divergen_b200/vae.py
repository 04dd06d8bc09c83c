"""`AutoencoderKL` -- decoder half only, the diffusers class surface over `dg_vae_*` (include/divergen_b200.h).

`StableDiffusionPipeline.__call__` step 8 is `vae.decode(latents / vae.config.scaling_factor).sample`; its result is what
the reference saves with `pt_to_pil(image)[j].save(out_path)` (DiverGen/generation/txt2img_diffusers_stages_from_txt.py:
255-267).  Keeps: `.config.{scaling_factor, latent_channels, block_out_channels, layers_per_block}`, `.dtype`, `.device`,
`.to()`, `.eval()`, `load_state_dict` with diffusers key names (encoder / quant_conv keys of a full `vae/` checkpoint are
accepted and ignored: generation never encodes), `decode(z, return_dict=True) -> .sample`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib


@dataclass
class DecoderOutput:
    sample: torch.Tensor


SD_VAE_CONFIG = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                     norm_num_groups=32, scaling_factor=0.18215)
_IGNORED_PREFIXES = ("encoder.", "quant_conv.")
# Published SD-1.x / 2.x `vae/diffusion_pytorch_model*.safetensors` files predate diffusers' Attention refactor: the mid-block
# attention is stored under the deprecated AttentionBlock names, which diffusers renames at load time
# (`_convert_deprecated_attention_blocks`).  Same mapping here.
_DEPRECATED_ATTN = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}


def _rename_deprecated(key: str) -> str:
    if ".attentions." in key:
        for old, new in _DEPRECATED_ATTN.items():
            if old in key:
                return key.replace(old, new)
    return key


class AutoencoderKL:
    def __init__(self, device="cuda:0", **config):
        cfg = dict(SD_VAE_CONFIG)
        cfg.update(config)
        if cfg["latent_channels"] != 4 or cfg["out_channels"] != 3 or cfg["norm_num_groups"] != 32:
            raise ValueError("only latent_channels=4, out_channels=3, norm_num_groups=32 are supported")
        if len(cfg["block_out_channels"]) != 4:
            raise ValueError("block_out_channels must have 4 entries")
        self.config = SimpleNamespace(**cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("divergen_b200 has no CPU path; device must be a CUDA (sm_100) device")
        self.dtype = torch.float16
        self._lib = _lib.load()
        self._ctx = _lib.context(self.device.index or 0)
        h = C.c_void_p()
        boc = (C.c_int32 * 4)(*cfg["block_out_channels"])
        _lib.check(self._lib.dg_vae_create(self._ctx, boc, cfg["layers_per_block"], C.byref(h)), "dg_vae_create")
        self._h = h
        self._prepared: Optional[Tuple[int, int, int]] = None

    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dg_vae_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def expected_state_dict_shapes(self) -> Dict[str, Tuple[int, ...]]:
        out = {}
        shp, nd = (C.c_int64 * 4)(), C.c_int32()
        for i in range(self._lib.dg_vae_num_weights(self._h)):
            name = self._lib.dg_vae_weight_name(self._h, i).decode()
            _lib.check(self._lib.dg_vae_weight_shape(self._h, i, shp, C.byref(nd)))
            out[name] = tuple(shp[k] for k in range(nd.value))
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        expected = self.expected_state_dict_shapes()
        state_dict = {_rename_deprecated(k): v for k, v in state_dict.items()}
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected and not k.startswith(_IGNORED_PREFIXES)]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} (+{max(0, len(missing) - 5)}), "
                               f"unexpected {unexpected[:5]} (+{max(0, len(unexpected) - 5)})")
        for k, v in state_dict.items():
            if k not in expected:
                continue
            t = v.detach().to(device=self.device, dtype=torch.float16).contiguous()
            if t.dim() == 4 and len(expected[k]) == 2:      # pre-0.18 checkpoints store the attention projections as 1x1 convs
                t = t.reshape(t.shape[0], t.shape[1])
            shp = (C.c_int64 * max(1, t.dim()))(*t.shape)
            _lib.check(self._lib.dg_vae_set_weight(self._h, k.encode(), C.c_void_p(t.data_ptr()), t.dim(), shp),
                       f"dg_vae_set_weight({k})")
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def prepare(self, max_batch: int, h: int, w: int):
        p = self._prepared
        if p is None or max_batch > p[0] or h * w > p[1] * p[2]:
            _lib.check(self._lib.dg_vae_prepare(self._h, max_batch, h, w), "dg_vae_prepare")
            self._prepared = (max_batch, h, w)

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None, scale: float = 1.0):
        """z: [B, 4, h, w] latents ALREADY divided by `config.scaling_factor` (the diffusers contract); `scale` lets the
        pipeline fold that division into the first kernel instead."""
        if z.dim() != 4 or z.shape[1] != self.config.latent_channels:
            raise ValueError(f"z must be [B, {self.config.latent_channels}, h, w], got {tuple(z.shape)}")
        if z.device != self.device:
            raise ValueError("z must live on the model's CUDA device")
        b, _, h, w = z.shape
        if h % 8 or w % 8:
            raise ValueError("latent height and width must be multiples of 8")
        if z.dtype != torch.float16 or not z.is_contiguous():
            z = z.to(torch.float16).contiguous()
        self.prepare(b, h, w)
        out = torch.empty((b, self.config.out_channels, 8 * h, 8 * w), dtype=torch.float16, device=self.device)
        stream = torch.cuda.current_stream(self.device)
        _lib.check(self._lib.dg_vae_decode(self._h, C.c_void_p(z.data_ptr()), float(scale), C.c_void_p(out.data_ptr()),
                                           b, h, w, C.c_void_p(stream.cuda_stream)), "dg_vae_decode")
        z.record_stream(stream)
        return DecoderOutput(sample=out) if return_dict else (out,)
