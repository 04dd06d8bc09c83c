"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

CPU restatement of diffusers' `DDIMScheduler` (eta = 0) and of the `StableDiffusionPipeline.__call__`
denoising loop, i.e. what runs inside `stage(...)` at
DiverGen/generation/txt2img_diffusers_stages_from_txt.py:255-259.  diffusers is absent, so the
constants follow the public SD scheduler config: scaled_linear betas
0.00085..0.012 over 1000 steps, clip_sample False, set_alpha_to_one False, steps_offset 1,
'leading' timestep spacing.  Known answers pinned in tests/test_oracle_ddim.py come from
SURVEY.md section 3.4.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class StepOutput:
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor


class DDIMOracle:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 prediction_type: str = "epsilon", steps_offset: int = 1, set_alpha_to_one: bool = False):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.steps_offset = steps_offset
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, eta: float = 0.0) -> StepOutput:
        assert eta == 0.0, "the oracle restates the deterministic (eta=0) DDIM step only"
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "v_prediction":
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            eps = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        else:
            raise ValueError(self.prediction_type)
        direction = (1 - a_prev) ** 0.5 * eps
        prev = a_prev ** 0.5 * x0 + direction
        return StepOutput(prev_sample=prev, pred_original_sample=x0)


def denoise_loop(unet, scheduler: DDIMOracle, latents: torch.Tensor, prompt_embeds: torch.Tensor,
                 negative_prompt_embeds: torch.Tensor, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, max_steps: Optional[int] = None) -> torch.Tensor:
    """Steps 4-7 of StableDiffusionPipeline.__call__.

    `max_steps` truncates the loop (bench cpu_baseline / tests time or check a bounded sample of it).
    """
    scheduler.set_timesteps(num_inference_steps)
    latents = latents * scheduler.init_noise_sigma
    ehs = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
    do_cfg = guidance_scale > 1.0
    for i, t in enumerate(scheduler.timesteps):
        if max_steps is not None and i >= max_steps:
            break
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        x_in = scheduler.scale_model_input(x_in, t)
        with torch.no_grad():
            noise = unet(x_in, t, ehs if do_cfg else prompt_embeds).sample
        if do_cfg:
            u, c = noise.chunk(2)
            noise = u + guidance_scale * (c - u)
        latents = scheduler.step(noise, t, latents).prev_sample
    return latents
