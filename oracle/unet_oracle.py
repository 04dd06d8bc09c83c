"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

Pure-PyTorch CPU restatement of diffusers' `UNet2DConditionModel.forward`, the class that
`DiffusionPipeline.from_pretrained` hands to the reference driver
(DiverGen/generation/txt2img_diffusers_stages_from_txt.py:139,182; called inside `pipe(...)` at
:255-259).  diffusers itself is absent from /root/reference and from this image, so the layer
forms follow the published architecture and the public SD-1.5 / SD-2.1
`unet/config.json` files.  Module / parameter names are the diffusers
state-dict names so a real checkpoint would load with `load_state_dict`.

Structural pins (tests/test_oracle_structure.py): 859 520 964 parameters / 686 tensors for SD-1.5,
865 910 724 for SD-2.1.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """Subset of diffusers' UNet2DConditionModel config that the SD-1.5 / SD-2.1 UNets use."""
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 64
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # diffusers' historic misnomer: `attention_head_dim` is the NUMBER OF HEADS per block.
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    upcast_attention: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    time_cond_proj_dim: Optional[int] = None
    # which down blocks carry attention (SD: first three)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    prediction_type: str = "epsilon"

    @staticmethod
    def sd15() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sd21() -> "UNetConfig":
        return UNetConfig(sample_size=96, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                          use_linear_projection=True, upcast_attention=True, prediction_type="v_prediction")

    @staticmethod
    def tiny(cross_attention_dim: int = 64, linear: bool = False) -> "UNetConfig":
        """A shrunken UNet with the same topology (for fast CPU/GPU parity tests)."""
        return UNetConfig(sample_size=16, block_out_channels=(64, 128, 128, 128), attention_head_dim=(2, 2, 4, 4),
                          cross_attention_dim=cross_attention_dim, use_linear_projection=linear)


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float,
                       max_period: int = 10000) -> torch.Tensor:
    """diffusers `get_timestep_embedding`: fp32 sinusoid, cos-first when flipped."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, cin: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: Optional[int], upcast: bool):
        super().__init__()
        self.heads = heads
        self.upcast = upcast
        ctx_dim = ctx_dim or dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, s, c = x.shape
        d = c // self.heads
        q = self.to_q(x).view(b, s, self.heads, d).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, self.heads, d).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, self.heads, d).transpose(1, 2)
        if self.upcast:
            q, k = q.float(), k.float()
        w = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1).to(v.dtype)
        o = (w @ v).transpose(1, 2).reshape(b, s, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)  # erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: int, upcast: bool):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, None, upcast)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim, upcast)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: int, groups: int, linear: bool, upcast: bool):
        super().__init__()
        self.linear = linear
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim, upcast)])
        self.proj_out = nn.Linear(dim, dim) if linear else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        r = x
        x = self.norm(x)
        if self.linear:
            x = self.proj_in(x.permute(0, 2, 3, 1).reshape(b, h * w, c))
        else:
            x = self.proj_in(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            x = blk(x, ctx)
        if self.linear:
            x = self.proj_out(x).reshape(b, h, w, c).permute(0, 3, 1, 2)
        else:
            x = self.proj_out(x.reshape(b, h, w, c).permute(0, 3, 1, 2))
        return x + r


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin: int, cout: int, heads: int, attn: bool, down: bool, temb: int):
        super().__init__()
        g = cfg.norm_num_groups
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, g, cfg.norm_eps)
                                      for i in range(cfg.layers_per_block)])
        if attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, heads, cfg.cross_attention_dim, g, cfg.use_linear_projection,
                                   cfg.upcast_attention) for _ in range(cfg.layers_per_block)])
        self.has_attn = attn
        if down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.has_down = down

    def forward(self, x, temb, ctx):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.has_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, c: int, heads: int, temb: int):
        super().__init__()
        g = cfg.norm_num_groups
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, cfg.cross_attention_dim, g,
                                                            cfg.use_linear_projection, cfg.upcast_attention)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, g, cfg.norm_eps) for _ in range(2)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin: int, cout: int, prev: int, heads: int, attn: bool, up: bool, temb: int):
        super().__init__()
        g = cfg.norm_num_groups
        n = cfg.layers_per_block + 1
        res = []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = prev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb, g, cfg.norm_eps))
        self.resnets = nn.ModuleList(res)
        if attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, heads, cfg.cross_attention_dim, g, cfg.use_linear_projection,
                                   cfg.upcast_attention) for _ in range(n)])
        self.has_attn = attn
        if up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.has_up = up

    def forward(self, x, skips: List[torch.Tensor], temb, ctx):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ctx)
        if self.has_up:
            x = self.upsamplers[0](x)
        return x


@dataclass
class UNetOutput:
    sample: torch.Tensor


class UNet2DConditionOracle(nn.Module):
    """Restatement of diffusers UNet2DConditionModel (SD-1.x / SD-2.x block layout)."""

    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        downs, cout = [], ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            downs.append(DownBlock(cfg, cin, cout, cfg.attention_head_dim[i], cfg.down_has_attn[i],
                                   i != len(ch) - 1, temb))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, ch[-1], cfg.attention_head_dim[-1], temb)
        rch = list(reversed(ch))
        rheads = list(reversed(cfg.attention_head_dim))
        rattn = list(reversed(cfg.down_has_attn))
        ups, cout = [], rch[0]
        for i, c in enumerate(rch):
            prev, cout = cout, c
            cin = rch[min(i + 1, len(ch) - 1)]
            ups.append(UpBlock(cfg, cin, cout, prev, rheads[i], rattn[i], i != len(ch) - 1, temb))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, encoder_hidden_states, return_dict: bool = True):
        cfg = self.cfg
        b = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.int64, device=sample.device)
        timestep = timestep.reshape(-1).expand(b) if timestep.numel() == 1 else timestep
        t_emb = timestep_embedding(timestep, cfg.block_out_channels[0], cfg.flip_sin_to_cos, cfg.freq_shift)
        emb = self.time_embedding(t_emb.to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return UNetOutput(sample=x) if return_dict else (x,)


def seeded_state_dict(cfg: UNetConfig, seed: int = 0, gain: float = 1.0):
    """Deterministic random-init weights (no checkpoints are available offline).

    Default torch init under a fixed seed; norm scales/biases are perturbed away from (1, 0) so the
    affine paths are exercised.  `gain` scales conv/linear weights to push fp16 range in tests.
    """
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = UNet2DConditionOracle(cfg)
    sd = model.state_dict()
    for k, v in sd.items():
        if ".norm" in k or "conv_norm_out" in k:
            if k.endswith("weight"):
                v.copy_(1.0 + 0.1 * torch.randn(v.shape, generator=g))
            else:
                v.copy_(0.05 * torch.randn(v.shape, generator=g))
        elif k.endswith("bias"):
            v.copy_(0.02 * torch.randn(v.shape, generator=g))
        elif gain != 1.0:
            v.mul_(gain)
    return sd
