"""Route to PINNED parity: regenerate the golden vectors from the REAL diffusers classes the reference calls.

    python tests/golden/make_golden_from_diffusers.py [--out tests/golden/diffusers_vectors.npz]

diffusers is absent from this image (and from /opt/wheelhouse; the reference neither vendors nor pins it), so today this script
exits with code 3 and `tests/test_diffusers_pin.py` is skipped: parity stays "unpinned".  The moment `import diffusers`
succeeds -- on any machine -- running it writes `diffusers_vectors.npz`, produced by upstream
`UNet2DConditionModel.forward`, `DDIMScheduler.step`, the `StableDiffusionPipeline.__call__` loop arithmetic and
`AutoencoderKL.decode` on the SAME seeded weights (`oracle.*.seeded_state_dict`: the oracle uses diffusers' state-dict key
names, so the weights load into the real classes unchanged) and the same seeded inputs as `make_golden.py`.
`tests/test_diffusers_pin.py` then checks the oracle against those vectors (CPU) and the CUDA path against them (GPU): from
that commit on the oracle is pinned to upstream and the header of oracle/__init__.py can say so.

Only public diffusers API is used: UNet2DConditionModel(**config), DDIMScheduler(**config), AutoencoderKL(**config).
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                set_alpha_to_one=False, steps_offset=1)


def unet_kwargs(cfg):
    """oracle UNetConfig -> diffusers UNet2DConditionModel constructor arguments (the published SD config keys)."""
    attn = list(cfg.down_has_attn) if hasattr(cfg, "down_has_attn") else [True, True, True, False]
    return dict(sample_size=cfg.sample_size, in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                layers_per_block=cfg.layers_per_block, block_out_channels=tuple(cfg.block_out_channels),
                down_block_types=tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in attn),
                up_block_types=tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(attn)),
                cross_attention_dim=cfg.cross_attention_dim, attention_head_dim=tuple(cfg.attention_head_dim),
                norm_num_groups=cfg.norm_num_groups, norm_eps=cfg.norm_eps, use_linear_projection=cfg.use_linear_projection,
                upcast_attention=cfg.upcast_attention, flip_sin_to_cos=True, freq_shift=0, act_fn="silu")


def generate(out_path):
    try:
        import diffusers
        from diffusers import AutoencoderKL, DDIMScheduler, UNet2DConditionModel
    except Exception as e:      # noqa: BLE001 -- any import failure means "not available here"
        print("diffusers is not importable here ({}): parity stays unpinned".format(e))
        return 3
    from oracle.unet_oracle import UNetConfig, seeded_state_dict
    from oracle.vae_oracle import VAEConfig, seeded_vae_state_dict
    torch.set_num_threads(1)
    out = {"diffusers_version": np.array(diffusers.__version__)}
    # ---- scheduler: set_timesteps + step on the make_golden.py inputs
    g = torch.Generator().manual_seed(7)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    out["step_x"], out["step_model_out"] = x.numpy(), e.numpy()
    for pred in ("epsilon", "v_prediction"):
        s = DDIMScheduler(prediction_type=pred, **SCHED_KW)
        s.set_timesteps(50)
        out["timesteps50"] = s.timesteps.numpy()
        out["alphas_cumprod"] = s.alphas_cumprod.numpy()
        for t in (981, 501, 1):
            out["step_{}_{}".format(pred, t)] = s.step(e, t, x, eta=0.0).prev_sample.numpy()
    # ---- UNet forward + 2-step CFG loop (the pipeline's loop arithmetic: cat, unet, chunk, guidance, scheduler.step)
    for name, linear in (("conv_proj", False), ("linear_proj", True)):
        cfg = UNetConfig.tiny(linear=linear)
        sd = {k: v.half().float() for k, v in seeded_state_dict(cfg, 0).items()}
        m = UNet2DConditionModel(**unet_kwargs(cfg)).eval()
        res = m.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        g = torch.Generator().manual_seed(42)
        lat = torch.randn(1, 4, 16, 16, generator=g).half().float()
        pos = torch.randn(1, 77, 64, generator=g).half().float()
        neg = torch.randn(1, 77, 64, generator=g).half().float()
        with torch.no_grad():
            out[name + "_forward_t981"] = m(torch.cat([lat, lat]), 981, encoder_hidden_states=torch.cat([neg, pos])).sample.numpy()
            s = DDIMScheduler(**SCHED_KW)
            s.set_timesteps(2)
            cur = lat * s.init_noise_sigma
            for t in s.timesteps:
                inp = s.scale_model_input(torch.cat([cur, cur]), t)
                u, c = m(inp, t, encoder_hidden_states=torch.cat([neg, pos])).sample.chunk(2)
                cur = s.step(u + 7.5 * (c - u), t, cur, eta=0.0).prev_sample
            out[name + "_loop2"] = cur.numpy()
        out[name + "_lat"], out[name + "_pos"], out[name + "_neg"] = lat.numpy(), pos.numpy(), neg.numpy()
    # ---- VAE decode
    vcfg = VAEConfig.tiny()
    vsd = {k: v.half().float() for k, v in seeded_vae_state_dict(vcfg, 0).items()}
    vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                        up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=tuple(vcfg.block_out_channels),
                        layers_per_block=vcfg.layers_per_block, latent_channels=4, norm_num_groups=32).eval()
    res = vae.load_state_dict(vsd, strict=False)
    assert not [k for k in res.missing_keys if k.startswith(("decoder.", "post_quant_conv."))], res.missing_keys
    assert not res.unexpected_keys, res.unexpected_keys
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(3)).half().float()
    with torch.no_grad():
        out["vae_z"], out["vae_decode"] = z.numpy(), vae.decode(z).sample.numpy()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes, diffusers", diffusers.__version__)
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "diffusers_vectors.npz"))
    raise SystemExit(generate(ap.parse_args().out))
