"""`DiffusionPipeline.from_pretrained(dir, variant='fp16', torch_dtype=torch.float16)` on a Hugging Face pipeline folder
written by tests/hf_layout.py (random weights, small config), used exactly as the reference script uses its pipelines
(txt2img_diffusers_stages_from_txt.py:139-143,242,255-267): .to(device), enable_model_cpu_offload, encode_prompt,
pipe(prompt_embeds=..., output_type='pt').images, pt_to_pil(...)[j].save(...)."""
import os

import pytest
import torch

from tests.test_gpu_unet import DEV, _need_gpu

pytestmark = pytest.mark.gpu


def test_from_pretrained_runs_the_reference_call_sequence(tmp_path):
    _need_gpu()
    from divergen_b200 import AutoencoderKL, CLIPTextModel, DDIMScheduler, DiffusionPipeline, UNet2DConditionModel
    from divergen_b200.utils import pt_to_pil
    from tests.hf_layout import write_pipeline_dir
    saved = write_pipeline_dir(str(tmp_path / "sd"), DEV)
    with pytest.warns(UserWarning):            # the SD-1.5 folder names PNDMScheduler
        stage = DiffusionPipeline.from_pretrained(os.path.join(str(tmp_path), "sd"), variant="fp16", torch_dtype=torch.float16)
        stage.to(torch.device(DEV))
    stage.enable_model_cpu_offload(0)
    stage.enable_xformers_memory_efficient_attention()
    assert isinstance(stage.unet, UNet2DConditionModel) and isinstance(stage.vae, AutoencoderKL)
    assert isinstance(stage.text_encoder, CLIPTextModel) and isinstance(stage.scheduler, DDIMScheduler)
    assert stage.safety_checker is None and stage.feature_extractor is None
    assert stage.unet.config.block_out_channels == (64, 128, 128, 128) and stage.vae_scale_factor == 8

    generator = torch.manual_seed(42)
    prompt_embeds, negative_embeds = stage.encode_prompt("a photo of a single aerosol can")
    assert prompt_embeds.shape == (1, 77, 64) and negative_embeds.shape == (1, 77, 64)
    image = stage(prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_embeds, generator=generator, output_type="pt",
                  num_images_per_prompt=2, num_inference_steps=3).images
    assert image.shape == (2, 3, 128, 128) and torch.isfinite(image.float()).all()
    out = tmp_path / "1_0000000.png"
    pt_to_pil(image)[0].save(out)
    from PIL import Image
    assert Image.open(out).size == (128, 128)

    # the loaded weights are the ones on disk: same result as components built by hand from the saved state dicts
    from divergen_b200 import StableDiffusionPipeline
    from divergen_b200.loading import unet_kwargs_from_config, vae_kwargs_from_config
    from tests.hf_layout import TINY_UNET_JSON, TINY_VAE_JSON
    unet = UNet2DConditionModel(device=DEV, **unet_kwargs_from_config(TINY_UNET_JSON))
    unet.load_state_dict(saved["unet"])
    vae = AutoencoderKL(device=DEV, **vae_kwargs_from_config(TINY_VAE_JSON))
    vae.load_state_dict(saved["vae"])
    pipe2 = StableDiffusionPipeline(unet, DDIMScheduler(), vae=vae)
    image2 = pipe2(prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_embeds, generator=torch.manual_seed(42),
                   output_type="pt", num_images_per_prompt=2, num_inference_steps=3).images
    # (not bit-equal: split-K partial sums are added in arrival order)
    assert (image.float() - image2.float()).abs().max().item() <= 2e-3


def test_from_pretrained_sd21_layout_and_overrides(tmp_path):
    """v-prediction scheduler config, no variant suffix, `text_encoder=None` override (reference :158-160)."""
    _need_gpu()
    from divergen_b200 import StableDiffusionPipeline
    from tests.hf_layout import SD21_SCHEDULER_JSON, write_pipeline_dir
    write_pipeline_dir(str(tmp_path / "sd21"), DEV, variant=None, scheduler_json=SD21_SCHEDULER_JSON, deprecated_vae_names=False)
    pipe = StableDiffusionPipeline.from_pretrained(str(tmp_path / "sd21"), text_encoder=None, variant="fp16", torch_dtype=torch.float16)
    pipe.to(DEV)
    assert pipe.text_encoder is None and pipe.tokenizer is None
    assert pipe.scheduler.config.prediction_type == "v_prediction"
    with pytest.raises(ValueError):
        pipe.encode_prompt("x")
    g = torch.Generator().manual_seed(0)
    pe, ne = torch.randn(1, 77, 64, generator=g).half(), torch.randn(1, 77, 64, generator=g).half()
    lat = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, output_type="latent", num_inference_steps=2).images
    assert lat.shape == (1, 4, 16, 16) and torch.isfinite(lat.float()).all()


def test_vae_accepts_deprecated_attention_names():
    """Published SD VAE checkpoints store the mid-block attention as query/key/value/proj_attn (1x1-conv or linear shapes)."""
    _need_gpu()
    from divergen_b200 import AutoencoderKL
    from tests.hf_layout import _rand_sd
    vae = AutoencoderKL(device=DEV, block_out_channels=(64, 64, 128, 128), layers_per_block=1)
    sd = _rand_sd(vae.expected_state_dict_shapes(), 5)
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(1)).half().to(DEV)
    vae.load_state_dict(sd)
    want = vae.decode(z).sample.clone()
    old = {}
    for k, v in sd.items():
        for new, dep in ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn.")):
            if ".attentions." in k and new in k:
                k = k.replace(new, dep)
                if k.endswith("weight"):
                    v = v[:, :, None, None]            # pre-0.18 files: 1x1 convolutions
        old[k] = v
    assert any(".query." in k for k in old)
    vae2 = AutoencoderKL(device=DEV, block_out_channels=(64, 64, 128, 128), layers_per_block=1)
    res = vae2.load_state_dict(old)
    assert not res.missing_keys and not res.unexpected_keys
    got = vae2.decode(z).sample
    assert (got.float() - want.float()).abs().max().item() <= 2e-3 * want.float().abs().max().item() + 1e-3
