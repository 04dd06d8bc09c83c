"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

CPU restatement of the decoder half of diffusers' `AutoencoderKL` as used by `StableDiffusionPipeline.__call__` step 8
(`image = vae.decode(latents / vae.config.scaling_factor).sample`): the step that turns
the denoised latents into the RGB image the reference saves at
DiverGen/generation/txt2img_diffusers_stages_from_txt.py:267.  diffusers is absent, so the structure follows the public
SD-1.x/2.x `vae/config.json` (block_out_channels 128/256/512/512, layers_per_block 2 -> 3 resnets per decoder up block,
norm_num_groups 32, eps 1e-6, one 512-wide single-head attention in the mid block, latent_channels 4,
scaling_factor 0.18215) with diffusers state-dict names.  Structural pin: 49 490 199 decoder-side parameters
(post_quant_conv + decoder), tests/test_oracle_structure.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215

    @staticmethod
    def sd() -> "VAEConfig":
        return VAEConfig()

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(64, 64, 128, 128))


class VaeResnet(nn.Module):
    def __init__(self, cin, cout, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):
    """diffusers `Attention` with `residual_connection=True`, one head of the full width, GroupNorm on the input."""
    def __init__(self, c, groups, eps):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax((q @ k.transpose(1, 2)) * (c ** -0.5), dim=-1)
        o = self.to_out[0](p @ v)
        return x + o.transpose(1, 2).reshape(b, c, h, w)


class VaeMid(nn.Module):
    def __init__(self, c, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(c, c, groups, eps), VaeResnet(c, c, groups, eps)])
        self.attentions = nn.ModuleList([VaeAttention(c, groups, eps)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VaeUp(nn.Module):
    def __init__(self, cin, cout, n, groups, eps, up):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(cin if i == 0 else cout, cout, groups, eps) for i in range(n)])
        self.upsamplers = nn.ModuleList([_Upsample(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class _Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch, g, eps = cfg.block_out_channels, cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[-1], 3, padding=1)
        self.mid_block = VaeMid(ch[-1], g, eps)
        rev = list(reversed(ch))
        ups, cin = [], rev[0]
        for i, cout in enumerate(rev):
            ups.append(VaeUp(cin, cout, cfg.layers_per_block + 1, g, eps, up=i != len(rev) - 1))
            cin = cout
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for u in self.up_blocks:
            x = u(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class VAEDecoderOracle(nn.Module):
    """`AutoencoderKL.decode`: post_quant_conv (1x1) -> Decoder.  Input: latents ALREADY divided by scaling_factor."""
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = _Decoder(cfg)

    def forward(self, z):
        return self.decoder(self.post_quant_conv(z))


def seeded_vae_state_dict(cfg: VAEConfig, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    sd = VAEDecoderOracle(cfg).state_dict()
    for k, v in sd.items():
        if "norm" in k:
            v.copy_(1.0 + 0.1 * torch.randn(v.shape, generator=g) if k.endswith("weight") else 0.05 * torch.randn(v.shape, generator=g))
        elif k.endswith("bias"):
            v.copy_(0.02 * torch.randn(v.shape, generator=g))
    return sd
