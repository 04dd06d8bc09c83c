"""`StableDiffusionPipeline` surface: `pipe(prompt_embeds=, negative_prompt_embeds=, generator=, output_type=,
num_images_per_prompt=, num_inference_steps=, guidance_scale=).images` as called by the reference driver
(DiverGen/generation/txt2img_diffusers_stages_from_txt.py:242,255-259), semantics per SURVEY.md 3.2.

In scope: latent preparation, the 50-step UNet + CFG + DDIM loop (one C-ABI call, CUDA-graph replayed).
Row f1 (VAE decode): pass `vae=AutoencoderKL(...)` (divergen_b200.vae, `dg_vae_decode`) or any `vae_decode` callable;
without either only `output_type='latent'` is available.  Row f2 (CLIP text encoder): pass
`text_encoder=CLIPTextModel(...)` (divergen_b200.clip, `dg_clip_encode`) together with a `tokenizer` (transformers
`CLIPTokenizer`; its vocabulary files are not shipped here), or any `text_encoder(prompt) -> embeddings` callable.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import torch

from .scheduler import DDIMScheduler
from .unet import UNet2DConditionModel


@dataclass
class StableDiffusionPipelineOutput:
    images: Union[torch.Tensor, list]
    nsfw_content_detected: Optional[List[bool]] = None


def pt_to_pil(images: torch.Tensor):
    """diffusers.utils.pt_to_pil: [B,3,H,W] in [-1,1] -> list of PIL images (txt2img_...py:267)."""
    from PIL import Image
    images = (images / 2 + 0.5).clamp(0, 1)
    arr = (images.cpu().permute(0, 2, 3, 1).float().numpy() * 255).round().astype("uint8")
    return [Image.fromarray(a) for a in arr]


_COMPONENTS = ("unet", "scheduler", "vae", "text_encoder", "tokenizer")


class StableDiffusionPipeline:
    def __init__(self, unet: UNet2DConditionModel, scheduler: DDIMScheduler,
                 text_encoder: Optional[Callable] = None, vae_decode: Optional[Callable] = None,
                 vae_scale_factor: int = 8, vae=None, tokenizer=None):
        self._lazy = None
        self.unet, self.scheduler = unet, scheduler
        self.tokenizer = tokenizer
        self.text_encoder, self.vae_decode = text_encoder, vae_decode
        self.vae = vae
        self.vae_scale_factor = vae_scale_factor
        self.device = unet.device
        self.feature_extractor = None
        self.safety_checker = None

    # ---- DiffusionPipeline.from_pretrained (divergen_b200/loading.py): the components are built on the GPU the caller names
    # with .to(device) / enable_model_cpu_offload(gpu_id) (reference :141-143), or on the current CUDA device at first use.
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        from .loading import DiffusionPipeline
        return DiffusionPipeline.from_pretrained(pretrained_model_name_or_path, **kwargs)

    def _init_lazy(self, spec):
        self._lazy = spec
        self.vae_decode = None
        self.vae_scale_factor = 8
        self.feature_extractor = None
        self.safety_checker = None

    def _materialize(self, device=None):
        if getattr(self, "_lazy", None) is None:
            return
        from .loading import materialize
        spec, self._lazy = self._lazy, None
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise ValueError("divergen_b200 has no CPU path; move the pipeline to a CUDA (sm_100) device")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        for k, v in materialize(spec, dev).items():
            setattr(self, k, v)
        self.device = dev
        if self.vae is not None:
            self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)

    def __getattr__(self, name):
        # only reached when normal lookup fails: a component of a not-yet-materialised from_pretrained pipeline
        if name in _COMPONENTS + ("device",) and self.__dict__.get("_lazy") is not None:
            self._materialize()
            return self.__dict__[name]
        raise AttributeError(name)

    @property
    def components(self):
        return {k: getattr(self, k) for k in _COMPONENTS}

    def to(self, device=None, *args, **kwargs):
        if isinstance(device, torch.dtype):               # pipe.to(torch.float16): the only dtype there is
            if device != torch.float16:
                raise ValueError("divergen_b200 stores fp16 only")
            return self
        if device is not None and getattr(self, "_lazy", None) is not None:
            self._materialize(device)
        elif device is not None and torch.device(device).type == "cuda":
            d = torch.device(device)
            if d.index is not None and d != self.device:
                raise ValueError("pipeline lives on {}; divergen_b200 components do not migrate between devices".format(self.device))
        elif device is not None:
            raise ValueError("divergen_b200 has no CPU path")
        return self

    def enable_model_cpu_offload(self, gpu_id=None, device=None):
        """Accepted for call compatibility (txt2img_...py:143); weights stay resident in HBM (1.7 GB of 180 GB)."""
        if getattr(self, "_lazy", None) is not None:
            self._materialize(device if device is not None else ("cuda:{}".format(gpu_id) if gpu_id is not None else None))

    def enable_xformers_memory_efficient_attention(self):
        """Accepted for call compatibility (txt2img_...py:186); attention is always the fused tcgen05 kernel."""

    def encode_prompt(self, prompt, device=None, num_images_per_prompt: int = 1, do_classifier_free_guidance: bool = True,
                      negative_prompt=None):
        if self.text_encoder is None:
            raise ValueError("no text_encoder attached: pass prompt_embeds/negative_prompt_embeds (CLIP is row f2, next)")
        neg_prompt = negative_prompt if negative_prompt is not None else ""
        if self.tokenizer is not None:
            # diffusers: text_encoder(tokenizer(prompt, padding="max_length", max_length=77, truncation=True).input_ids)[0]
            def enc(text):
                ids = self.tokenizer(text, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                                     return_tensors="pt").input_ids
                return self.text_encoder(ids)[0]
            return enc(prompt), enc(neg_prompt)
        return self.text_encoder(prompt), self.text_encoder(neg_prompt)

    @torch.no_grad()
    def __call__(self, prompt=None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                 num_images_per_prompt: int = 1, eta: float = 0.0, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 output_type: str = "pil", return_dict: bool = True, cross_attention_kwargs=None):
        if eta != 0.0:
            raise ValueError("only eta=0 (deterministic DDIM) is supported")
        if cross_attention_kwargs is not None:
            raise ValueError("cross_attention_kwargs is not supported")
        cfg = self.unet.config
        height = height or cfg.sample_size * self.vae_scale_factor
        width = width or cfg.sample_size * self.vae_scale_factor
        if height % 64 or width % 64:
            raise ValueError("height and width must be multiples of 64")
        do_cfg = guidance_scale > 1.0
        if prompt_embeds is None:
            if prompt is None:
                raise ValueError("provide `prompt` or `prompt_embeds`")
            prompt_embeds, negative_prompt_embeds = self.encode_prompt(prompt, negative_prompt=negative_prompt)
        if do_cfg and negative_prompt_embeds is None:
            raise ValueError("classifier-free guidance needs negative_prompt_embeds")
        dev = self.device

        def h2d(t):
            # Host tensors in PINNED memory are copied without blocking the host (the copy is ordered on the stream like every
            # kernel of this call), so a driver that keeps its inputs pinned enqueues call k+1 while the GPU still runs call k --
            # a blocking copy would wait for the previous call to drain and expose this call's launch work (~20 ms of an idle
            # GPU per call, measured in the C5 slice).  As with any non_blocking copy the caller must not overwrite a pinned
            # input before the stream has consumed it; pageable inputs are staged synchronously, exactly as `.to(device)`.
            return t.to(dev, torch.float16, non_blocking=bool(t.device.type == "cpu" and t.is_pinned()))

        pe = h2d(prompt_embeds)
        bsz = pe.shape[0] * num_images_per_prompt
        pe = pe.repeat_interleave(num_images_per_prompt, dim=0)
        if do_cfg:
            ne = h2d(negative_prompt_embeds).repeat_interleave(num_images_per_prompt, dim=0)
            ehs = torch.cat([ne, pe], dim=0).contiguous()
        else:
            ehs = pe.contiguous()
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        shape = (bsz, cfg.in_channels, h, w)
        if latents is None:
            # diffusers randn_tensor: a CPU generator draws on the CPU and moves (the reference passes
            # torch.manual_seed(seed + rank), a CPU generator: txt2img_...py:200)
            gdev = generator.device if generator is not None else torch.device("cpu")
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=torch.float16).to(dev)
        else:
            latents = h2d(latents)
        latents = (latents * self.scheduler.init_noise_sigma).contiguous().clone()
        self.scheduler.set_timesteps(num_inference_steps)
        ts = [int(t) for t in self.scheduler.timesteps]
        al = [self.scheduler.alphas_for(t) for t in ts]
        self.unet.denoise_loop(latents, ehs, ts, [a for a, _ in al], [p for _, p in al], guidance_scale,
                               self.scheduler.config.prediction_type)
        if output_type == "latent":
            images = latents
        else:
            if self.vae is not None:
                # image = vae.decode(latents / vae.config.scaling_factor).sample, the division folded into the first kernel
                images = self.vae.decode(latents, scale=1.0 / self.vae.config.scaling_factor).sample
            elif self.vae_decode is not None:
                images = self.vae_decode(latents / 0.18215)
            else:
                raise ValueError("output_type != 'latent' needs `vae=` (divergen_b200.AutoencoderKL) or a vae_decode callable")
            if output_type == "pt":
                images = (images / 2 + 0.5).clamp(0, 1)
            elif output_type in ("pil", "uint8"):
                # pt_to_pil's arithmetic on the device (row f4): [B, H, W, 3] uint8; "uint8" hands that tensor to the caller
                # (the driver copies it to pinned memory asynchronously and PNG-encodes on worker threads)
                from . import ops
                from PIL import Image
                images = ops.image_to_uint8(images.to(torch.float16).contiguous())
                if output_type == "pil":
                    images = [Image.fromarray(a) for a in images.cpu().numpy()]
            else:
                raise ValueError(output_type)
        return StableDiffusionPipelineOutput(images=images) if return_dict else (images, None)
