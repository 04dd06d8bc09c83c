"""`UNet2DConditionModel` -- the diffusers class surface over the C-ABI library.

Keeps: `.config.{in_channels, sample_size, time_cond_proj_dim, cross_attention_dim}`, `.dtype`, `.device`, `.to()`,
`.eval()`, `load_state_dict` with diffusers key names (686 tensors for SD-1.5 / SD-2.1) and
`forward(sample, timestep, encoder_hidden_states, ..., return_dict=True) -> .sample`.  Unsupported optional
arguments (class labels, ControlNet residuals, attention masks, LoRA scale) raise instead of being ignored.
The reference reaches this class through `DiffusionPipeline.from_pretrained(...)` / `pipe(...)` at
DiverGen/generation/txt2img_diffusers_stages_from_txt.py:139,255-259.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import _lib


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor


SD15_CONFIG = dict(in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280),
                   layers_per_block=2, attention_head_dim=(8, 8, 8, 8), cross_attention_dim=768, norm_num_groups=32,
                   norm_eps=1e-5, use_linear_projection=False, upcast_attention=False, flip_sin_to_cos=True,
                   freq_shift=0, time_cond_proj_dim=None, down_has_attn=(True, True, True, False))
SD21_CONFIG = dict(SD15_CONFIG, sample_size=96, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                   use_linear_projection=True, upcast_attention=True)


class UNet2DConditionModel:
    def __init__(self, device="cuda:0", **config):
        cfg = dict(SD15_CONFIG)
        cfg.update(config)
        heads = cfg["attention_head_dim"]
        if isinstance(heads, int):
            heads = (heads,) * 4
        cfg["attention_head_dim"] = tuple(heads)
        if cfg.get("time_cond_proj_dim") is not None:
            raise ValueError("time_cond_proj_dim is not supported")
        self.config = SimpleNamespace(**cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("divergen_b200 has no CPU path; device must be a CUDA (sm_100) device")
        self.dtype = torch.float16
        self._lib = _lib.load()
        self._ctx = _lib.context(self.device.index or 0)
        c = _lib.UNetConfigC()
        c.in_channels, c.out_channels, c.sample_size = cfg["in_channels"], cfg["out_channels"], cfg["sample_size"]
        c.block_out_channels = (C.c_int32 * 4)(*cfg["block_out_channels"])
        c.layers_per_block = cfg["layers_per_block"]
        c.num_heads = (C.c_int32 * 4)(*heads)
        c.cross_attention_dim, c.norm_num_groups = cfg["cross_attention_dim"], cfg["norm_num_groups"]
        c.norm_eps = cfg["norm_eps"]
        c.use_linear_projection, c.upcast_attention = int(cfg["use_linear_projection"]), int(cfg["upcast_attention"])
        c.down_has_attn = (C.c_int32 * 4)(*[int(b) for b in cfg["down_has_attn"]])
        c.flip_sin_to_cos, c.freq_shift = int(cfg["flip_sin_to_cos"]), float(cfg["freq_shift"])
        h = C.c_void_p()
        _lib.check(self._lib.dg_unet_create(self._ctx, C.byref(c), C.byref(h)), "dg_unet_create")
        self._h = h
        self._prepared: Optional[Tuple[int, int, int, int]] = None

    # ---- torch.nn.Module-ish surface the pipeline / driver touches
    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.dg_unet_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def expected_state_dict_shapes(self) -> Dict[str, Tuple[int, ...]]:
        n = self._lib.dg_unet_num_weights(self._h)
        out = {}
        shp, nd = (C.c_int64 * 4)(), C.c_int32()
        for i in range(n):
            name = self._lib.dg_unet_weight_name(self._h, i).decode()
            _lib.check(self._lib.dg_unet_weight_shape(self._h, i, shp, C.byref(nd)))
            out[name] = tuple(shp[k] for k in range(nd.value))
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        expected = self.expected_state_dict_shapes()
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} (+{max(0, len(missing) - 5)}), "
                               f"unexpected {unexpected[:5]} (+{max(0, len(unexpected) - 5)})")
        for k, v in state_dict.items():
            if k not in expected:
                continue
            t = v.detach().to(device=self.device, dtype=torch.float16).contiguous()
            shp = (C.c_int64 * max(1, t.dim()))(*t.shape)
            _lib.check(self._lib.dg_unet_set_weight(self._h, k.encode(), C.c_void_p(t.data_ptr()), t.dim(), shp),
                       f"dg_unet_set_weight({k})")
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def prepare(self, max_batch: int, h: int, w: int, ctx_tokens: int = 77):
        key = (max_batch, h, w, ctx_tokens)
        p = self._prepared
        if p is None or max_batch > p[0] or h * w > p[1] * p[2] or ctx_tokens > p[3]:
            _lib.check(self._lib.dg_unet_prepare(self._h, max_batch, h, w, ctx_tokens), "dg_unet_prepare")
            self._prepared = key

    def set_graphs(self, enabled: bool):
        _lib.check(self._lib.dg_unet_set_graphs(self._h, int(enabled)))

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.dg_unet_last_launch_count(self._h))

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                encoder_attention_mask=None, return_dict: bool = True, out: Optional[torch.Tensor] = None):
        for name, val in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                          ("attention_mask", attention_mask), ("cross_attention_kwargs", cross_attention_kwargs),
                          ("added_cond_kwargs", added_cond_kwargs),
                          ("down_block_additional_residuals", down_block_additional_residuals),
                          ("mid_block_additional_residual", mid_block_additional_residual),
                          ("encoder_attention_mask", encoder_attention_mask)):
            if val is not None:
                raise ValueError(f"UNet2DConditionModel.forward: `{name}` is not supported by divergen_b200")
        if sample.dim() != 4 or sample.shape[1] != self.config.in_channels:
            raise ValueError(f"sample must be [B, {self.config.in_channels}, h, w], got {tuple(sample.shape)}")
        if sample.device != self.device or encoder_hidden_states.device != self.device:
            raise ValueError("sample / encoder_hidden_states must live on the model's CUDA device")
        b, _, h, w = sample.shape
        ehs = encoder_hidden_states
        if ehs.dim() != 3 or ehs.shape[0] != b or ehs.shape[2] != self.config.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states must be [B, tokens, {self.config.cross_attention_dim}]")
        if sample.dtype != torch.float16 or not sample.is_contiguous():
            sample = sample.to(torch.float16).contiguous()
        if ehs.dtype != torch.float16 or not ehs.is_contiguous():
            ehs = ehs.to(torch.float16).contiguous()
        if torch.is_tensor(timestep):
            ts = timestep.detach().reshape(-1).float().cpu().tolist()
        else:
            ts = [float(timestep)]
        if len(ts) not in (1, b):
            raise ValueError("timestep must be a scalar or have one entry per sample")
        self.prepare(b, h, w, ehs.shape[1])
        if out is None:
            out = torch.empty((b, self.config.out_channels, h, w), dtype=torch.float16, device=self.device)
        tarr = (C.c_float * len(ts))(*ts)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.dg_unet_forward(self._h, C.c_void_p(sample.data_ptr()), tarr, len(ts),
                                             C.c_void_p(ehs.data_ptr()), ehs.shape[1], C.c_void_p(out.data_ptr()),
                                             b, h, w, C.c_void_p(stream)), "dg_unet_forward")
        # keep inputs alive until the asynchronous work is enqueued behind them on this stream
        sample.record_stream(torch.cuda.current_stream(self.device))
        ehs.record_stream(torch.cuda.current_stream(self.device))
        return UNet2DConditionOutput(sample=out) if return_dict else (out,)

    __call__ = forward

    def denoise_loop(self, latents, ehs, timesteps, alpha_t, alpha_prev, guidance_scale: float, prediction_type: str):
        """dg_denoise_loop: all steps of StableDiffusionPipeline.__call__'s loop on the device (latents in place)."""
        n, _, h, w = latents.shape
        b = 2 * n if guidance_scale > 1.0 else n
        assert latents.dtype == torch.float16 and latents.is_contiguous() and ehs.dtype == torch.float16 and ehs.is_contiguous()
        assert ehs.shape[0] == b
        self.prepare(b, h, w, ehs.shape[1])
        k = len(timesteps)
        fa = lambda xs: (C.c_float * k)(*[float(x) for x in xs])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.dg_denoise_loop(self._h, C.c_void_p(latents.data_ptr()), C.c_void_p(ehs.data_ptr()),
                                             ehs.shape[1], n, h, w, fa(timesteps), fa(alpha_t), fa(alpha_prev), k,
                                             float(guidance_scale), {"epsilon": 0, "v_prediction": 1}[prediction_type],
                                             C.c_void_p(stream)), "dg_denoise_loop")
        return latents
