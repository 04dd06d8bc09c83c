"""GPU parity of `AutoencoderKL.decode` through the diffusers-shaped surface / `dg_vae_decode`,
against oracle/vae_oracle.py on identical seeded weights and latents.  Same statement as tests/test_gpu_unet.py:
max|err| <= 2e-2 * max|ref| + 1e-3, cosine >= 0.999, rms(E_new) <= 2 * rms(E_torch) + 1e-4 where E_torch is the oracle
run in fp16 eager on the GPU against the oracle in fp32 on the CPU.
"""
import pytest
import torch

from tests.test_gpu_unet import DEV, _check, _need_gpu

pytestmark = pytest.mark.gpu


def _models(cfg, seed=0):
    from divergen_b200 import AutoencoderKL
    from oracle.vae_oracle import VAEDecoderOracle, seeded_vae_state_dict
    sd = {k: v.half().float() for k, v in seeded_vae_state_dict(cfg, seed).items()}
    oracle = VAEDecoderOracle(cfg).eval()
    oracle.load_state_dict(sd)
    vae = AutoencoderKL(device=DEV, block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block)
    res = vae.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    return oracle, vae


@pytest.mark.parametrize("B,h,w", [(1, 8, 8), (3, 16, 8), (2, 16, 16)])
def test_tiny_vae_decode(B, h, w):
    _need_gpu()
    from oracle.vae_oracle import VAEConfig
    oracle, vae = _models(VAEConfig.tiny())
    z = torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(3)).half()
    with torch.no_grad():
        ref32 = oracle(z.float())
        ref16 = oracle.half().to(DEV)(z.to(DEV))
    got = vae.decode(z.to(DEV)).sample
    assert got.shape == (B, 3, 8 * h, 8 * w)
    _check(got, ref32, ref16, name=f"vae tiny B={B} {h}x{w}")


def test_tiny_vae_scale_folding_and_independence():
    """decode(z, scale=s) == decode(z * s); a sample's image does not depend on what else is in the batch."""
    _need_gpu()
    from oracle.vae_oracle import VAEConfig
    _, vae = _models(VAEConfig.tiny())
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(4)).half().to(DEV)
    s = 1.0 / vae.config.scaling_factor
    a = vae.decode(z, scale=s).sample
    b = vae.decode((z.float() * s).half()).sample
    assert (a.float() - b.float()).abs().max().item() <= 2e-2 * b.float().abs().max().item() + 1e-3
    one = vae.decode(z[1:2].contiguous(), scale=s).sample
    # not bit-equal: the split-K choice of a GEMM depends on the row count, and split-K sums in arrival order
    assert (one[0].float() - a[1].float()).abs().max().item() <= 5e-3 * a.float().abs().max().item() + 1e-3


def test_sd_vae_decode_full_width():
    """The SD VAE (128/256/512/512, 49.5 M parameters) on one 32x32 latent -> 256x256 image."""
    _need_gpu()
    from oracle.vae_oracle import VAEConfig
    oracle, vae = _models(VAEConfig.sd())
    assert len(vae.expected_state_dict_shapes()) == 140
    z = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(5)).half()
    with torch.no_grad():
        ref32 = oracle(z.float())
        ref16 = oracle.half().to(DEV)(z.to(DEV))
    got = vae.decode(z.to(DEV)).sample
    _check(got, ref32, ref16, name="vae sd 32x32")


def test_vae_errors():
    _need_gpu()
    from divergen_b200 import AutoencoderKL
    vae = AutoencoderKL(device=DEV, block_out_channels=(64, 64, 128, 128))
    with pytest.raises(RuntimeError):          # weights not set
        vae.decode(torch.zeros(1, 4, 8, 8, dtype=torch.float16, device=DEV))
    with pytest.raises(ValueError):
        vae.decode(torch.zeros(1, 3, 8, 8, dtype=torch.float16, device=DEV))
    with pytest.raises(ValueError):
        vae.decode(torch.zeros(1, 4, 12, 8, dtype=torch.float16, device=DEV))
    with pytest.raises(RuntimeError):
        vae.load_state_dict({"decoder.bogus": torch.zeros(1)})


def test_pipeline_with_vae_pt_output():
    """pipe(..., output_type='pt').images in [0, 1], [N, 3, 8h, 8w] -- the tensor the reference hands to pt_to_pil."""
    _need_gpu()
    from divergen_b200 import DDIMScheduler, StableDiffusionPipeline
    from oracle.unet_oracle import UNetConfig
    from oracle.vae_oracle import VAEConfig
    from tests.test_gpu_unet import _models as unet_models
    ucfg = UNetConfig.tiny()
    _, unet = unet_models(ucfg)
    _, vae = _models(VAEConfig.tiny())
    pipe = StableDiffusionPipeline(unet, DDIMScheduler(), vae=vae)
    g = torch.Generator().manual_seed(0)
    pe = torch.randn(2, 77, ucfg.cross_attention_dim, generator=g)
    ne = torch.randn(2, 77, ucfg.cross_attention_dim, generator=g)
    img = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, generator=torch.manual_seed(7), output_type="pt",
               num_inference_steps=4, height=128, width=128).images
    assert img.shape == (2, 3, 128, 128) and torch.isfinite(img).all()
    assert img.min().item() >= 0.0 and img.max().item() <= 1.0


def test_sd_vae_decode_full_size_512():
    """BASELINE's image size: one 64x64 latent -> 512x512 through the SD VAE (the 4096-token mid-block attention and the
    262144-pixel 128-channel convolutions), against the fp32 oracle on the host."""
    _need_gpu()
    from oracle.vae_oracle import VAEConfig
    oracle, vae = _models(VAEConfig.sd(), seed=2)
    z = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(9)).half()
    with torch.no_grad():
        ref32 = oracle(z.float())
        ref16 = oracle.half().to(DEV)(z.to(DEV))
    got = vae.decode(z.to(DEV)).sample
    assert got.shape == (1, 3, 512, 512)
    _check(got, ref32, ref16, name="vae sd 64x64 -> 512x512")
