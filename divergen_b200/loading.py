"""`DiffusionPipeline.from_pretrained` -- the loader half of the diffusers surface.

The reference builds its pipelines with
    DiffusionPipeline.from_pretrained(os.path.join(args.ckpt_dir, <name>), variant='fp16', torch_dtype=torch.float16)
    stage.to(device); stage.enable_model_cpu_offload(local_rank)
(DiverGen/generation/txt2img_diffusers_stages_from_txt.py:139-143,182-187).  This module reads the same on-disk layout --
a Hugging Face pipeline directory:

    model_index.json                      {"_class_name": "StableDiffusionPipeline", "unet": [...], "vae": [...], ...}
    unet/config.json                      + unet/diffusion_pytorch_model[.fp16].safetensors
    vae/config.json                       + vae/diffusion_pytorch_model[.fp16].safetensors
    text_encoder/config.json              + text_encoder/model[.fp16].safetensors
    tokenizer/{vocab.json,merges.txt,...}   (transformers CLIPTokenizer)
    scheduler/scheduler_config.json

and returns a `StableDiffusionPipeline` whose components are this package's classes.  Host side only: configuration parsing
and file IO; the weights go straight into the packed device layout through `load_state_dict` (C-ABI `dg_*_set_weight`).

Device placement follows the reference's call order: `from_pretrained` only records what to load (there is no CPU model in
this package), `.to(device)` / `enable_model_cpu_offload(gpu_id)` / the first call materialise the components on the GPU.
"""
from __future__ import annotations

import json
import os
import warnings
from typing import Optional

import torch

_UNSET = object()

# unet/config.json keys whose non-default values change the arithmetic and are not built here
_UNET_MUST_BE = {
    "act_fn": ("silu", "swish"), "center_input_sample": (False,), "class_embed_type": (None,), "num_class_embeds": (None,),
    "addition_embed_type": (None,), "dual_cross_attention": (False,), "resnet_time_scale_shift": ("default",),
    "time_embedding_type": ("positional",), "timestep_post_act": (None,), "conv_in_kernel": (3,), "conv_out_kernel": (3,),
    "mid_block_type": ("UNetMidBlock2DCrossAttn",), "encoder_hid_dim": (None,), "time_cond_proj_dim": (None,),
    "class_embeddings_concat": (False,), "mid_block_only_cross_attention": (None, False), "cross_attention_norm": (None,),
    "transformer_layers_per_block": (1,), "downsample_padding": (1,), "mid_block_scale_factor": (1, 1.0),
    "resnet_skip_time_act": (False,), "resnet_out_scale_factor": (1, 1.0), "time_embedding_act_fn": (None,),
    "time_embedding_dim": (None,), "projection_class_embeddings_input_dim": (None,), "attention_type": ("default",),
}


def _read_json(path: str) -> dict:
    with open(path, "r") as f:
        return json.load(f)


def unet_kwargs_from_config(cfg: dict) -> dict:
    """unet/config.json (diffusers UNet2DConditionModel config) -> keyword arguments of divergen_b200.UNet2DConditionModel.
    Anything this library does not build raises instead of being ignored."""
    for k, allowed in _UNET_MUST_BE.items():
        if k in cfg and cfg[k] not in allowed:
            raise ValueError("unet/config.json: {}={!r} is not supported by divergen_b200 (supported: {})".format(k, cfg[k], allowed))
    down = list(cfg.get("down_block_types", ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)))
    up = list(cfg.get("up_block_types", ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3))
    if len(down) != 4 or len(up) != 4:
        raise ValueError("unet/config.json: exactly 4 down / up blocks are supported")
    for t in down:
        if t not in ("CrossAttnDownBlock2D", "DownBlock2D"):
            raise ValueError("unet/config.json: down block type {!r} is not supported".format(t))
    has_attn = tuple(t == "CrossAttnDownBlock2D" for t in down)
    want_up = ["CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(has_attn)]
    if up != want_up:
        raise ValueError("unet/config.json: up_block_types {} do not mirror down_block_types {}".format(up, down))
    oca = cfg.get("only_cross_attention", False)
    if oca not in (False, None) and any(oca if isinstance(oca, (list, tuple)) else [oca]):
        raise ValueError("unet/config.json: only_cross_attention is not supported")
    if cfg.get("num_attention_heads") is not None:
        raise ValueError("unet/config.json: num_attention_heads is not supported (Stable Diffusion configs use attention_head_dim)")
    # NB: in the SD-1.x / 2.x configs `attention_head_dim` is (despite its name) the NUMBER OF HEADS per block
    heads = cfg.get("attention_head_dim", 8)
    heads = tuple(heads) if isinstance(heads, (list, tuple)) else (int(heads),) * 4
    return dict(in_channels=cfg.get("in_channels", 4), out_channels=cfg.get("out_channels", 4),
                sample_size=cfg.get("sample_size", 64), block_out_channels=tuple(cfg.get("block_out_channels", (320, 640, 1280, 1280))),
                layers_per_block=cfg.get("layers_per_block", 2), attention_head_dim=heads,
                cross_attention_dim=cfg.get("cross_attention_dim", 1280), norm_num_groups=cfg.get("norm_num_groups", 32),
                norm_eps=cfg.get("norm_eps", 1e-5), use_linear_projection=bool(cfg.get("use_linear_projection", False)),
                upcast_attention=bool(cfg.get("upcast_attention", False)), flip_sin_to_cos=bool(cfg.get("flip_sin_to_cos", True)),
                freq_shift=cfg.get("freq_shift", 0), down_has_attn=has_attn)


def vae_kwargs_from_config(cfg: dict) -> dict:
    if cfg.get("act_fn", "silu") not in ("silu", "swish"):
        raise ValueError("vae/config.json: act_fn={!r} is not supported".format(cfg.get("act_fn")))
    return dict(latent_channels=cfg.get("latent_channels", 4), out_channels=cfg.get("out_channels", 3),
                block_out_channels=tuple(cfg.get("block_out_channels", (128, 256, 512, 512))),
                layers_per_block=cfg.get("layers_per_block", 2), norm_num_groups=cfg.get("norm_num_groups", 32),
                scaling_factor=cfg.get("scaling_factor", 0.18215))


def text_encoder_kwargs_from_config(cfg: dict) -> dict:
    """text_encoder/config.json (transformers CLIPTextConfig; older files nest it under `text_config`)."""
    c = cfg.get("text_config") or cfg
    return dict(vocab_size=c.get("vocab_size", 49408), hidden_size=c.get("hidden_size", 768),
                intermediate_size=c.get("intermediate_size", 3072), num_hidden_layers=c.get("num_hidden_layers", 12),
                num_attention_heads=c.get("num_attention_heads", 12), max_position_embeddings=c.get("max_position_embeddings", 77),
                hidden_act=c.get("hidden_act", "quick_gelu"))


def scheduler_from_config(cfg: dict):
    """scheduler/scheduler_config.json -> DDIMScheduler.  SD-1.x checkpoints name PNDMScheduler there; this library builds
    the DDIM sampler the task's hot path names (`DDIMScheduler.step`), so any other class is re-read as DDIM over the same
    noise schedule -- what `DDIMScheduler.from_config(pipe.scheduler.config)` does in diffusers -- and said so once."""
    from .scheduler import DDIMScheduler
    name = cfg.get("_class_name", "DDIMScheduler")
    if name != "DDIMScheduler":
        warnings.warn("scheduler_config.json names {}; divergen_b200 samples with DDIMScheduler (eta 0) over the same "
                      "beta schedule".format(name), stacklevel=3)
    return DDIMScheduler.from_config(cfg)


def find_weights(folder: str, stem: str, variant: Optional[str]) -> str:
    """`<stem>.<variant>.safetensors` first, as diffusers resolves `variant='fp16'`, then the plain name, then `.bin`."""
    names = []
    if variant:
        names += ["{}.{}.safetensors".format(stem, variant), "{}.{}.bin".format(stem, variant)]
    names += ["{}.safetensors".format(stem), "{}.bin".format(stem)]
    for n in names:
        p = os.path.join(folder, n)
        if os.path.exists(p):
            return p
    raise FileNotFoundError("no weights in {} (looked for {})".format(folder, ", ".join(names)))


def read_state_dict(path: str) -> dict:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


class DiffusionPipeline:
    """`diffusers.DiffusionPipeline`: only the loader entry point; the object returned is a StableDiffusionPipeline."""

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, variant: Optional[str] = None, torch_dtype=None,
                        text_encoder=_UNSET, tokenizer=_UNSET, vae=_UNSET, unet=_UNSET, scheduler=_UNSET,
                        safety_checker=None, feature_extractor=None, watermarker=None, requires_safety_checker: bool = False,
                        use_safetensors: Optional[bool] = None, local_files_only: bool = True, **unused):
        from .pipeline import StableDiffusionPipeline
        path = str(pretrained_model_name_or_path)
        if not os.path.isdir(path):
            raise FileNotFoundError("{} is not a directory: divergen_b200 loads local pipeline folders only (no hub access)".format(path))
        if torch_dtype not in (None, torch.float16):
            raise ValueError("torch_dtype={} is not supported: the sm_100a path stores fp16 and accumulates in fp32".format(torch_dtype))
        if unused:
            raise TypeError("from_pretrained: unsupported arguments {}".format(sorted(unused)))
        index = _read_json(os.path.join(path, "model_index.json"))
        klass = index.get("_class_name", "StableDiffusionPipeline")
        if klass != "StableDiffusionPipeline":
            raise ValueError("model_index.json names {}; divergen_b200 builds StableDiffusionPipeline only".format(klass))
        overrides = {"text_encoder": text_encoder, "tokenizer": tokenizer, "vae": vae, "unet": unet, "scheduler": scheduler}
        spec = {"path": path, "variant": variant, "overrides": {k: v for k, v in overrides.items() if v is not _UNSET}}
        pipe = StableDiffusionPipeline.__new__(StableDiffusionPipeline)
        pipe._init_lazy(spec)
        pipe.safety_checker, pipe.feature_extractor = safety_checker, feature_extractor
        return pipe


def materialize(spec: dict, device: torch.device) -> dict:
    """Build the components named by the pipeline folder on `device` and load their weights."""
    from . import AutoencoderKL, CLIPTextModel, UNet2DConditionModel
    path, variant, ov = spec["path"], spec["variant"], spec["overrides"]
    out = {}
    if "unet" in ov:
        out["unet"] = ov["unet"]
    else:
        d = os.path.join(path, "unet")
        m = UNet2DConditionModel(device=device, **unet_kwargs_from_config(_read_json(os.path.join(d, "config.json"))))
        m.load_state_dict(read_state_dict(find_weights(d, "diffusion_pytorch_model", variant)))
        out["unet"] = m
    if "scheduler" in ov:
        out["scheduler"] = ov["scheduler"]
    else:
        out["scheduler"] = scheduler_from_config(_read_json(os.path.join(path, "scheduler", "scheduler_config.json")))
    if "vae" in ov:
        out["vae"] = ov["vae"]
    elif os.path.isdir(os.path.join(path, "vae")):
        d = os.path.join(path, "vae")
        m = AutoencoderKL(device=device, **vae_kwargs_from_config(_read_json(os.path.join(d, "config.json"))))
        m.load_state_dict(read_state_dict(find_weights(d, "diffusion_pytorch_model", variant)))
        out["vae"] = m
    else:
        out["vae"] = None
    if "text_encoder" in ov:
        out["text_encoder"] = ov["text_encoder"]           # e.g. text_encoder=None (reference :158-160)
    elif os.path.isdir(os.path.join(path, "text_encoder")):
        d = os.path.join(path, "text_encoder")
        m = CLIPTextModel(device=device, **text_encoder_kwargs_from_config(_read_json(os.path.join(d, "config.json"))))
        m.load_state_dict(read_state_dict(find_weights(d, "model", variant)))
        out["text_encoder"] = m
    else:
        out["text_encoder"] = None
    if "tokenizer" in ov:
        out["tokenizer"] = ov["tokenizer"]
    elif out["text_encoder"] is not None and os.path.isdir(os.path.join(path, "tokenizer")):
        from transformers import CLIPTokenizer
        out["tokenizer"] = CLIPTokenizer.from_pretrained(os.path.join(path, "tokenizer"))
    else:
        out["tokenizer"] = None
    return out
