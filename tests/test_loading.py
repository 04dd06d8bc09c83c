"""Host-side checks of the `DiffusionPipeline.from_pretrained` loader (divergen_b200/loading.py): configuration parsing of
the published SD-1.5 / SD-2.1 `unet/config.json` contents, refusal of configurations the library does not build, weight-file
resolution for `variant='fp16'`.  (The reference's call: txt2img_diffusers_stages_from_txt.py:139-143.)"""
import json
import os

import pytest
import torch

from divergen_b200 import SD15_CONFIG, SD21_CONFIG
from divergen_b200.loading import (DiffusionPipeline, find_weights, scheduler_from_config, text_encoder_kwargs_from_config,
                                   unet_kwargs_from_config)
from tests.hf_layout import SD15_SCHEDULER_JSON, SD21_SCHEDULER_JSON

# unet/config.json as published (runwayml/stable-diffusion-v1-5, stabilityai/stable-diffusion-2-1)
SD15_UNET_JSON = {"_class_name": "UNet2DConditionModel", "_diffusers_version": "0.6.0", "act_fn": "silu", "attention_head_dim": 8,
                  "block_out_channels": [320, 640, 1280, 1280], "center_input_sample": False, "cross_attention_dim": 768,
                  "down_block_types": ["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
                  "downsample_padding": 1, "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": 4, "layers_per_block": 2,
                  "mid_block_scale_factor": 1, "norm_eps": 1e-05, "norm_num_groups": 32, "out_channels": 4, "sample_size": 64,
                  "up_block_types": ["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"]}
SD21_UNET_JSON = dict(SD15_UNET_JSON, _diffusers_version="0.10.0.dev0", attention_head_dim=[5, 10, 20, 20], cross_attention_dim=1024,
                      dual_cross_attention=False, num_class_embeds=None, only_cross_attention=False, sample_size=96,
                      upcast_attention=True, use_linear_projection=True)


def test_published_unet_configs_map_to_the_builtin_ones():
    for js, want in ((SD15_UNET_JSON, SD15_CONFIG), (SD21_UNET_JSON, SD21_CONFIG)):
        got = unet_kwargs_from_config(js)
        for k, v in got.items():
            assert want[k] == v or tuple(want[k]) == tuple(v), (k, want[k], v)


@pytest.mark.parametrize("key,val", [("class_embed_type", "timestep"), ("addition_embed_type", "text_time"), ("act_fn", "gelu"),
                                     ("transformer_layers_per_block", 2), ("only_cross_attention", True),
                                     ("num_attention_heads", 8), ("resnet_time_scale_shift", "scale_shift")])
def test_unsupported_unet_config_raises(key, val):
    with pytest.raises(ValueError):
        unet_kwargs_from_config(dict(SD15_UNET_JSON, **{key: val}))
    with pytest.raises(ValueError):
        unet_kwargs_from_config(dict(SD15_UNET_JSON, down_block_types=["DownBlock2D", "AttnDownBlock2D", "DownBlock2D", "DownBlock2D"]))


def test_scheduler_configs():
    with pytest.warns(UserWarning):
        s15 = scheduler_from_config(SD15_SCHEDULER_JSON)          # PNDM config -> DDIM over the same schedule, announced
    assert s15.config.prediction_type == "epsilon" and s15.config.steps_offset == 1 and not s15.config.set_alpha_to_one
    s21 = scheduler_from_config(SD21_SCHEDULER_JSON)
    assert s21.config.prediction_type == "v_prediction"
    s21.set_timesteps(50)
    assert int(s21.timesteps[0]) == 981 and int(s21.timesteps[-1]) == 1
    with pytest.raises(ValueError):
        scheduler_from_config(dict(SD21_SCHEDULER_JSON, clip_sample=True))


def test_text_encoder_config_sd21():
    js = {"hidden_act": "gelu", "hidden_size": 1024, "intermediate_size": 4096, "num_attention_heads": 16, "num_hidden_layers": 23,
          "max_position_embeddings": 77, "vocab_size": 49408, "projection_dim": 512}
    from divergen_b200 import SD21_CLIP_CONFIG
    assert text_encoder_kwargs_from_config(js) == SD21_CLIP_CONFIG


def test_variant_resolution(tmp_path):
    d = tmp_path / "unet"
    d.mkdir()
    (d / "diffusion_pytorch_model.safetensors").write_bytes(b"")
    assert find_weights(str(d), "diffusion_pytorch_model", "fp16").endswith("diffusion_pytorch_model.safetensors")
    (d / "diffusion_pytorch_model.fp16.safetensors").write_bytes(b"")
    assert find_weights(str(d), "diffusion_pytorch_model", "fp16").endswith(".fp16.safetensors")
    assert find_weights(str(d), "diffusion_pytorch_model", None).endswith("diffusion_pytorch_model.safetensors")
    with pytest.raises(FileNotFoundError):
        find_weights(str(d), "model", "fp16")


def test_from_pretrained_argument_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        DiffusionPipeline.from_pretrained(str(tmp_path / "nope"))
    (tmp_path / "model_index.json").write_text(json.dumps({"_class_name": "IFPipeline"}))
    with pytest.raises(ValueError):
        DiffusionPipeline.from_pretrained(str(tmp_path), variant="fp16", torch_dtype=torch.float16)
    (tmp_path / "model_index.json").write_text(json.dumps({"_class_name": "StableDiffusionPipeline"}))
    with pytest.raises(ValueError):
        DiffusionPipeline.from_pretrained(str(tmp_path), torch_dtype=torch.float32)
    pipe = DiffusionPipeline.from_pretrained(str(tmp_path), variant="fp16", torch_dtype=torch.float16)   # lazy: nothing touched yet
    with pytest.raises(ValueError):
        pipe.to("cpu")


def test_reference_import_lines():
    """The two import lines of the reference script (:7-8), pointed at this package."""
    from divergen_b200 import DiffusionPipeline as DP
    from divergen_b200.utils import pt_to_pil
    assert DP is DiffusionPipeline and callable(pt_to_pil)
