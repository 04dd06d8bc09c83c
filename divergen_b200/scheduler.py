"""`DDIMScheduler` surface (set_timesteps, timesteps, init_noise_sigma, scale_model_input, step, order, config).

Host-side schedule arithmetic (fp32 `alphas_cumprod`, 'leading' spacing, steps_offset) in numpy; the tensor update of
`step` is the fused sm_100a kernel behind `dg_cfg_ddim_step` (guidance disabled -> plain DDIM step).
Replaces diffusers DDIMScheduler as used inside `pipe(...)` (txt2img_diffusers_stages_from_txt.py:255-259).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import _lib


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", clip_sample: bool = False, set_alpha_to_one: bool = False,
                 steps_offset: int = 1, prediction_type: str = "epsilon", timestep_spacing: str = "leading"):
        if beta_schedule != "scaled_linear" or clip_sample or timestep_spacing != "leading":
            raise ValueError("only the Stable-Diffusion DDIM config (scaled_linear, no clipping, leading) is supported")
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(prediction_type)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, clip_sample=clip_sample,
                                      set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset,
                                      prediction_type=prediction_type, timestep_spacing=timestep_spacing)
        betas = np.linspace(np.float32(beta_start) ** 0.5, np.float32(beta_end) ** 0.5, num_train_timesteps,
                            dtype=np.float32) ** 2
        self.alphas_cumprod = np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)
        self.final_alpha_cumprod = np.float32(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **overrides):
        """`DDIMScheduler.from_config(scheduler_config.json dict | another scheduler's .config)`: the noise-schedule fields
        are taken, sampler-specific fields of other scheduler classes (PNDM's `skip_prk_steps`, ...) are dropped.
        `clip_sample` defaults to False here (the Stable-Diffusion DDIM config); a config that asks for clipping raises."""
        c = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        c.update(overrides)
        if c.get("trained_betas") is not None:
            raise ValueError("trained_betas is not supported")
        if c.get("thresholding") or c.get("rescale_betas_zero_snr"):
            raise ValueError("thresholding / rescale_betas_zero_snr are not supported")
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "clip_sample", "set_alpha_to_one",
                "steps_offset", "prediction_type", "timestep_spacing")
        return cls(**{k: c[k] for k in keys if k in c})

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)  # kept on the host: the loop only needs their values

    def scale_model_input(self, sample, timestep=None):
        return sample

    def alphas_for(self, timestep: int):
        prev_t = int(timestep) - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[int(timestep)])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("call set_timesteps first")
        if eta != 0.0 or variance_noise is not None:
            raise ValueError("only the deterministic eta=0 DDIM step is supported")
        if model_output.device.type != "cuda":
            raise ValueError("divergen_b200 has no CPU path")
        a_t, a_prev = self.alphas_for(int(timestep))
        prev = sample.to(torch.float16).contiguous().clone()
        noise = model_output.to(torch.float16).contiguous()
        n = prev.shape[0]
        ctx = _lib.context(prev.device.index or 0)
        stream = torch.cuda.current_stream(prev.device).cuda_stream
        _lib.check(_lib.load().dg_cfg_ddim_step(ctx, C.c_void_p(noise.data_ptr()), C.c_void_p(prev.data_ptr()), n,
                                                prev.numel() // n, a_t, a_prev, 1.0,
                                                0 if self.config.prediction_type == "epsilon" else 1,
                                                C.c_void_p(stream)), "dg_cfg_ddim_step")
        noise.record_stream(torch.cuda.current_stream(prev.device))
        return DDIMSchedulerOutput(prev_sample=prev) if return_dict else (prev,)
