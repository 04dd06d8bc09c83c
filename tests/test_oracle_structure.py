"""Structural pins of the oracle (the reference holds no golden vectors for this path: SURVEY.md 4, 8c)."""
import torch

from oracle.unet_oracle import UNet2DConditionOracle, UNetConfig, seeded_state_dict, timestep_embedding


def _meta(cfg):
    with torch.device("meta"):
        return UNet2DConditionOracle(cfg)


def test_sd15_param_count_and_keys():
    m = _meta(UNetConfig.sd15())
    assert sum(p.numel() for p in m.parameters()) == 859_520_964
    sd = m.state_dict()
    assert len(sd) == 686
    for k in ("time_embedding.linear_1.weight", "down_blocks.0.resnets.0.norm1.weight",
              "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight",
              "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight",
              "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.2.weight",
              "down_blocks.0.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.1.upsamplers.0.conv.weight", "conv_norm_out.weight", "conv_out.weight"):
        assert k in sd, k
    assert tuple(sd["down_blocks.0.attentions.0.proj_in.weight"].shape) == (320, 320, 1, 1)
    assert tuple(sd["up_blocks.1.resnets.2.conv1.weight"].shape) == (1280, 1920, 3, 3)


def test_sd21_param_count():
    m = _meta(UNetConfig.sd21())
    assert sum(p.numel() for p in m.parameters()) == 865_910_724
    sd = m.state_dict()
    assert len(sd) == 686
    assert tuple(sd["down_blocks.0.attentions.0.proj_in.weight"].shape) == (320, 320)
    assert tuple(sd["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"].shape) == (320, 1024)


def test_timestep_embedding_known_answer():
    # t = 981, dim 320, cos-first
    t = timestep_embedding(torch.tensor([981]), 320, True, 0)[0]
    assert torch.allclose(t[0:3], torch.tensor([0.67995721, -0.79842919, 0.57806414]), atol=1e-5)
    assert torch.allclose(t[160:163], torch.tensor([0.73325181, 0.60208869, 0.81599128]), atol=1e-5)
    assert abs(t[319].item() - 0.10372588) < 1e-5


def test_config1_plumbing_cpu():
    """BASELINE config 1: single 64x64 latent (CFG pair), 1 DDIM step, random-init UNet on CPU -- tiny-width variant
    of the same topology keeps the CPU suite fast; the full-width case is timed by bench.py's cpu_baseline."""
    from oracle.ddim_oracle import DDIMOracle, denoise_loop
    torch.manual_seed(42)
    cfg = UNetConfig.tiny()
    m = UNet2DConditionOracle(cfg).eval()
    m.load_state_dict(seeded_state_dict(cfg, 0))
    lat = torch.randn(1, 4, 16, 16)
    pos, neg = torch.randn(1, 77, cfg.cross_attention_dim), torch.randn(1, 77, cfg.cross_attention_dim)
    out = denoise_loop(m, DDIMOracle(), lat, pos, neg, num_inference_steps=50, guidance_scale=7.5, max_steps=1)
    assert out.shape == lat.shape and torch.isfinite(out).all()
    assert not torch.allclose(out, lat)


def test_vae_decoder_param_count_and_keys():
    """SD-1.x/2.x `vae/` checkpoint, decoder side (post_quant_conv + decoder): 49 490 199 parameters in 140 tensors."""
    from oracle.vae_oracle import VAEConfig, VAEDecoderOracle
    with torch.device("meta"):
        m = VAEDecoderOracle(VAEConfig.sd())
    assert sum(p.numel() for p in m.parameters()) == 49_490_199
    sd = m.state_dict()
    assert len(sd) == 140
    for k in ("post_quant_conv.weight", "decoder.conv_in.weight", "decoder.mid_block.attentions.0.to_q.weight",
              "decoder.mid_block.attentions.0.to_out.0.bias", "decoder.mid_block.resnets.1.conv2.weight",
              "decoder.up_blocks.0.upsamplers.0.conv.weight", "decoder.up_blocks.2.resnets.0.conv_shortcut.weight",
              "decoder.up_blocks.3.resnets.2.norm2.bias", "decoder.conv_norm_out.weight", "decoder.conv_out.bias"):
        assert k in sd, k
    assert tuple(sd["decoder.up_blocks.2.resnets.0.conv1.weight"].shape) == (256, 512, 3, 3)
    assert tuple(sd["decoder.up_blocks.3.resnets.0.conv1.weight"].shape) == (128, 256, 3, 3)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd


def test_vae_decoder_tiny_cpu():
    from oracle.vae_oracle import VAEConfig, VAEDecoderOracle, seeded_vae_state_dict
    cfg = VAEConfig.tiny()
    m = VAEDecoderOracle(cfg).eval()
    m.load_state_dict(seeded_vae_state_dict(cfg, 0))
    with torch.no_grad():
        y = m(torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(1)))
    assert y.shape == (2, 3, 64, 64) and torch.isfinite(y).all()


def test_clip_text_model_upstream_structure():
    """Row f2's checker is transformers.CLIPTextModel itself: pin the SD-1.x text tower's size and key names."""
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu")
    with torch.device("meta"):
        m = CLIPTextModel(cfg)
    assert sum(p.numel() for p in m.parameters()) == 123_060_480
    keys = [k for k in m.state_dict() if not k.endswith("position_ids")]
    assert len(keys) == 196
    for k in ("text_model.embeddings.token_embedding.weight", "text_model.embeddings.position_embedding.weight",
              "text_model.encoder.layers.0.self_attn.q_proj.bias", "text_model.encoder.layers.11.mlp.fc2.weight",
              "text_model.final_layer_norm.weight"):
        assert k in keys, k
