"""Test helper: write a Hugging Face Stable-Diffusion pipeline directory (the layout `DiffusionPipeline.from_pretrained`
reads: model_index.json, unet/, vae/, text_encoder/, tokenizer/, scheduler/) with seeded random weights of a small config.
No checkpoint is reachable offline, so the layout -- file names, `variant='fp16'` suffixes, config keys, the deprecated VAE
attention key names of the published checkpoints -- is what these fixtures pin, not the weight values."""
import json
import os

import torch

TINY_UNET_JSON = {
    "_class_name": "UNet2DConditionModel", "_diffusers_version": "0.6.0", "act_fn": "silu", "attention_head_dim": 2,
    "block_out_channels": [64, 128, 128, 128], "center_input_sample": False, "cross_attention_dim": 64,
    "down_block_types": ["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    "downsample_padding": 1, "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": 4, "layers_per_block": 1,
    "mid_block_scale_factor": 1, "norm_eps": 1e-05, "norm_num_groups": 32, "out_channels": 4, "sample_size": 16,
    "up_block_types": ["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"]}
TINY_VAE_JSON = {
    "_class_name": "AutoencoderKL", "act_fn": "silu", "block_out_channels": [64, 64, 128, 128], "in_channels": 3,
    "latent_channels": 4, "layers_per_block": 1, "norm_num_groups": 32, "out_channels": 3, "sample_size": 128,
    "scaling_factor": 0.18215, "down_block_types": ["DownEncoderBlock2D"] * 4, "up_block_types": ["UpDecoderBlock2D"] * 4}
TINY_TEXT_JSON = {
    "architectures": ["CLIPTextModel"], "hidden_act": "quick_gelu", "hidden_size": 64, "intermediate_size": 128,
    "max_position_embeddings": 77, "model_type": "clip_text_model", "num_attention_heads": 1, "num_hidden_layers": 2,
    "vocab_size": 514, "layer_norm_eps": 1e-05}
# scheduler_config.json of runwayml/stable-diffusion-v1-5 (PNDM) and of stabilityai/stable-diffusion-2-1 (DDIM, v-prediction)
SD15_SCHEDULER_JSON = {"_class_name": "PNDMScheduler", "_diffusers_version": "0.6.0", "beta_end": 0.012,
                       "beta_schedule": "scaled_linear", "beta_start": 0.00085, "num_train_timesteps": 1000,
                       "set_alpha_to_one": False, "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None,
                       "clip_sample": False}
SD21_SCHEDULER_JSON = {"_class_name": "DDIMScheduler", "_diffusers_version": "0.8.0", "beta_end": 0.012,
                       "beta_schedule": "scaled_linear", "beta_start": 0.00085, "clip_sample": False,
                       "num_train_timesteps": 1000, "prediction_type": "v_prediction", "set_alpha_to_one": False,
                       "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None}


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return [chr(c) for c in cs]


def write_tokenizer(d):
    """A byte-level CLIP vocabulary without merges (514 entries): loads with transformers.CLIPTokenizer."""
    os.makedirs(d, exist_ok=True)
    toks = _bytes_to_unicode()
    vocab = toks + [t + "</w>" for t in toks] + ["<|startoftext|>", "<|endoftext|>"]
    json.dump({t: i for i, t in enumerate(vocab)}, open(os.path.join(d, "vocab.json"), "w"))
    open(os.path.join(d, "merges.txt"), "w").write("#version: 0.2\n")
    json.dump({"model_max_length": 77, "bos_token": "<|startoftext|>", "eos_token": "<|endoftext|>",
               "unk_token": "<|endoftext|>", "pad_token": "<|endoftext|>", "tokenizer_class": "CLIPTokenizer"},
              open(os.path.join(d, "tokenizer_config.json"), "w"))


def _rand_sd(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        fan_in = 1
        for x in shp[1:]:
            fan_in *= x
        if "norm" in k and k.endswith("weight"):
            sd[k] = (1.0 + 0.05 * torch.randn(shp, generator=g)).half()
        elif k.endswith("bias"):
            sd[k] = (0.02 * torch.randn(shp, generator=g)).half()
        else:
            sd[k] = (torch.randn(shp, generator=g) / max(1.0, fan_in) ** 0.5).half()
    return sd


def write_pipeline_dir(root, device, variant="fp16", scheduler_json=None, with_text_encoder=True, deprecated_vae_names=True,
                       seed=0):
    """Write the directory and return the state dicts that were saved ({'unet': ..., 'vae': ..., 'text_encoder': ...})."""
    from safetensors.torch import save_file
    from divergen_b200 import AutoencoderKL, CLIPTextModel, UNet2DConditionModel
    from divergen_b200.loading import text_encoder_kwargs_from_config, unet_kwargs_from_config, vae_kwargs_from_config
    suffix = ".{}".format(variant) if variant else ""
    os.makedirs(root, exist_ok=True)
    index = {"_class_name": "StableDiffusionPipeline", "_diffusers_version": "0.6.0",
             "feature_extractor": [None, None], "safety_checker": [None, None], "requires_safety_checker": False,
             "scheduler": ["diffusers", (scheduler_json or SD15_SCHEDULER_JSON)["_class_name"]],
             "tokenizer": ["transformers", "CLIPTokenizer"], "unet": ["diffusers", "UNet2DConditionModel"],
             "vae": ["diffusers", "AutoencoderKL"]}
    if with_text_encoder:
        index["text_encoder"] = ["transformers", "CLIPTextModel"]
    json.dump(index, open(os.path.join(root, "model_index.json"), "w"))
    saved = {}
    os.makedirs(os.path.join(root, "unet"), exist_ok=True)
    json.dump(TINY_UNET_JSON, open(os.path.join(root, "unet", "config.json"), "w"))
    unet = UNet2DConditionModel(device=device, **unet_kwargs_from_config(TINY_UNET_JSON))
    saved["unet"] = _rand_sd(unet.expected_state_dict_shapes(), seed)
    save_file(saved["unet"], os.path.join(root, "unet", "diffusion_pytorch_model{}.safetensors".format(suffix)))
    os.makedirs(os.path.join(root, "vae"), exist_ok=True)
    json.dump(TINY_VAE_JSON, open(os.path.join(root, "vae", "config.json"), "w"))
    vae = AutoencoderKL(device=device, **vae_kwargs_from_config(TINY_VAE_JSON))
    vsd = _rand_sd(vae.expected_state_dict_shapes(), seed + 1)
    saved["vae"] = dict(vsd)
    if deprecated_vae_names:      # as in the published checkpoints: query/key/value/proj_attn
        ren = {".to_q.": ".query.", ".to_k.": ".key.", ".to_v.": ".value.", ".to_out.0.": ".proj_attn."}
        out = {}
        for k, v in vsd.items():
            for new, old in ren.items():
                if ".attentions." in k and new in k:
                    k = k.replace(new, old)
            out[k] = v
        vsd = out
    vsd["encoder.conv_in.weight"] = torch.zeros(64, 3, 3, 3).half()          # a full VAE checkpoint also carries the encoder
    vsd["quant_conv.weight"] = torch.zeros(8, 8, 1, 1).half()
    save_file(vsd, os.path.join(root, "vae", "diffusion_pytorch_model{}.safetensors".format(suffix)))
    os.makedirs(os.path.join(root, "scheduler"), exist_ok=True)
    json.dump(scheduler_json or SD15_SCHEDULER_JSON, open(os.path.join(root, "scheduler", "scheduler_config.json"), "w"))
    write_tokenizer(os.path.join(root, "tokenizer"))
    if with_text_encoder:
        os.makedirs(os.path.join(root, "text_encoder"), exist_ok=True)
        json.dump(TINY_TEXT_JSON, open(os.path.join(root, "text_encoder", "config.json"), "w"))
        te = CLIPTextModel(device=device, **text_encoder_kwargs_from_config(TINY_TEXT_JSON))
        saved["text_encoder"] = _rand_sd(te.expected_state_dict_shapes(), seed + 2)
        tsd = dict(saved["text_encoder"])
        tsd["text_model.embeddings.position_ids"] = torch.arange(77)[None]      # old transformers checkpoints carry this buffer
        save_file(tsd, os.path.join(root, "text_encoder", "model{}.safetensors".format(suffix)))
    return saved
