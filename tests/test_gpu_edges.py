"""Edge cases of the hot path on the GPU: ragged / non-square / odd shapes, context lengths other than 77, guidance off,
images-per-prompt > 1, seeding.  Same tolerances as tests/test_gpu_unet.py (fp16 storage, fp32 accumulate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


def _models(seed=0, linear=False):
    from tests.test_gpu_unet import _models as mk
    from oracle.unet_oracle import UNetConfig
    return mk(UNetConfig.tiny(cross_attention_dim=64, linear=linear), seed)


def _close(got, ref, name, rel=2e-2, abs_=1e-3):
    got, ref = got.float().cpu(), ref.float().cpu()
    e = (got - ref).abs().max().item()
    lim = rel * ref.abs().max().item() + abs_
    print(f"{name}: max err {e:.4g} (limit {lim:.4g})")
    assert torch.isfinite(got).all() and e <= lim, name


@pytest.mark.parametrize("B,h,w,tokens", [(1, 16, 24, 77), (5, 8, 8, 77), (2, 24, 16, 1), (3, 16, 16, 100), (2, 40, 8, 13)])
def test_unet_ragged_shapes(B, h, w, tokens):
    """Non-square latents, odd batches (partial M tiles, samples beyond the last tile), context lengths 1 / 13 / 100."""
    _need_gpu()
    oracle, unet = _models(seed=2)
    g = torch.Generator().manual_seed(100 + B + h + tokens)
    x = torch.randn(B, 4, h, w, generator=g).half()
    ehs = torch.randn(B, tokens, 64, generator=g).half()
    with torch.no_grad():
        ref = oracle(x.float(), 501, ehs.float()).sample
    got = unet(x.to(DEV), 501, ehs.to(DEV)).sample
    torch.cuda.synchronize()
    _close(got, ref, f"unet B={B} {h}x{w} tokens={tokens}")


def test_unet_rejects_bad_latent_size():
    _need_gpu()
    _, unet = _models(seed=1)
    x = torch.zeros(1, 4, 12, 16, device=DEV, dtype=torch.half)
    ehs = torch.zeros(1, 77, 64, device=DEV, dtype=torch.half)
    with pytest.raises((ValueError, RuntimeError)):
        unet(x, 1, ehs)          # 12 is not a multiple of 8: three stride-2 levels would not round-trip


@pytest.mark.parametrize("Sq,Sk", [(128, 1), (200, 77), (300, 129), (64, 100), (1, 77), (257, 255)])
def test_attention_ragged_lengths(Sq, Sk):
    _need_gpu()
    from divergen_b200 import ops
    from tests.test_gpu_ops import _attn_ref, _rand
    B, heads, d = 2, 4, 40
    q = _rand(B, Sq, heads * d, seed=70)
    kv = _rand(B, Sk, 2 * heads * d, seed=71)
    k, v = kv[..., :heads * d], kv[..., heads * d:]
    got = ops.attention(q, k, v, heads)
    ref = _attn_ref(q, k, v, heads)
    _close(got, ref, f"attention {Sq}x{Sk}", rel=1e-2, abs_=3e-3)


def test_ddim_step_every_timestep_of_the_schedule():
    """All 50 steps of the schedule (including the last, which uses final_alpha_cumprod): <= 2 fp16 ulp vs the oracle."""
    _need_gpu()
    from divergen_b200 import DDIMScheduler
    from oracle.ddim_oracle import DDIMOracle
    for pred in ("epsilon", "v_prediction"):
        s, o = DDIMScheduler(prediction_type=pred), DDIMOracle(prediction_type=pred)
        s.set_timesteps(50)
        o.set_timesteps(50)
        g = torch.Generator().manual_seed(3)
        x = torch.randn(2, 4, 16, 16, generator=g).half()
        e = torch.randn(2, 4, 16, 16, generator=g).half()
        for t in s.timesteps.tolist():
            got = s.step(e.to(DEV), t, x.to(DEV)).prev_sample.float().cpu()
            ref = o.step(e.float(), t, x.float()).prev_sample
            ulp = torch.clamp(ref.abs(), min=2.0 ** -14) * 2.0 ** -10
            assert ((got - ref).abs() <= 2 * ulp + 1e-6).all(), (pred, t)


def test_pipeline_guidance_off_images_per_prompt_and_seeding():
    _need_gpu()
    from divergen_b200 import DDIMScheduler, StableDiffusionPipeline
    from oracle.ddim_oracle import DDIMOracle, denoise_loop
    oracle, unet = _models(seed=4)
    pipe = StableDiffusionPipeline(unet, DDIMScheduler())
    g = torch.Generator().manual_seed(9)
    pos, neg = torch.randn(1, 77, 64, generator=g).half(), torch.randn(1, 77, 64, generator=g).half()
    lat = torch.randn(1, 4, 16, 16, generator=g).half()
    # guidance_scale = 1: no CFG batch doubling, negative embeddings unused
    out = pipe(prompt_embeds=pos, latents=lat, num_inference_steps=3, guidance_scale=1.0, height=128, width=128,
               output_type="latent").images
    ref = denoise_loop(oracle, DDIMOracle(), lat.float(), pos.float(), neg.float(), num_inference_steps=3, guidance_scale=1.0)
    _close(out, ref, "loop without guidance", rel=3e-2, abs_=2e-3)
    # num_images_per_prompt: latents drawn from the CPU generator the reference passes (torch.manual_seed(seed + rank))
    a = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, generator=torch.manual_seed(42), num_images_per_prompt=3,
             num_inference_steps=2, height=128, width=128, output_type="latent").images
    b = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, generator=torch.manual_seed(42), num_images_per_prompt=3,
             num_inference_steps=2, height=128, width=128, output_type="latent").images
    c = pipe(prompt_embeds=pos, negative_prompt_embeds=neg, generator=torch.manual_seed(43), num_images_per_prompt=3,
             num_inference_steps=2, height=128, width=128, output_type="latent").images
    assert a.shape == (3, 4, 16, 16)
    assert (a.float() - b.float()).abs().max().item() <= 2e-2 * a.float().abs().max().item()    # same seed -> same images (split-K fp32 order aside)
    assert (a.float() - c.float()).abs().max().item() > 0.1                                     # seed + 1 (another rank) -> different images
    assert (a[0].float() - a[1].float()).abs().max().item() > 0.1                                # images of one call differ
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pos, negative_prompt_embeds=neg, eta=0.5, num_inference_steps=2, output_type="latent")
    with pytest.raises(ValueError):
        pipe(prompt_embeds=pos, negative_prompt_embeds=neg, num_inference_steps=2, height=100, width=128, output_type="latent")


def test_image_to_uint8_matches_pt_to_pil():
    """Row f4: the device-side uint8 conversion is bit-identical to diffusers' pt_to_pil arithmetic (as restated in
    divergen_b200.pipeline.pt_to_pil) on values inside, outside and exactly on the clamp / rounding boundaries."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import numpy as np
    from divergen_b200 import ops, pt_to_pil
    g = torch.Generator().manual_seed(0)
    img = (torch.randn(3, 3, 40, 24, generator=g) * 0.8).half()
    img[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, -1.5, 1.5, 1 / 255, -1 / 255, 0.00392]).half()
    got = ops.image_to_uint8(img.cuda()).cpu().numpy()
    want = np.stack([np.asarray(p) for p in pt_to_pil(img.cuda())])
    assert got.shape == (3, 40, 24, 3) and got.dtype == np.uint8
    assert np.array_equal(got, want)
