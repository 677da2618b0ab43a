"""Oracle restatement of the AutoencoderKL DECODER behind ``vae.decode`` (test infrastructure only; torch fp32).

The reference calls it right after the denoise loop (src/pipelines/pipeline_diffsensei.py:339-367):
``latents / scaling_factor -> vae.decode -> image_processor.postprocess``.  ``AutoencoderKL`` lives in the
un-vendored, un-pinned ``diffusers`` dependency, so this restates its published SDXL-VAE semantics:
**parity unpinned** (tests/test_oracle_diffusers_pin.py pins it whenever diffusers is importable).

    decode(z)      = Decoder(post_quant_conv(z))                       post_quant_conv: Conv2d(4, 4, 1)
    Decoder        = conv_in 3x3 (4 -> C3) -> UNetMidBlock2D -> 4 x UpDecoderBlock2D -> GroupNorm(32, eps 1e-6)
                     -> SiLU -> conv_out 3x3 (C0 -> 3)                 block_out_channels (C0..C3) = (128, 256, 512, 512)
    UNetMidBlock2D = ResnetBlock2D, Attention (1 head of width C3 over all H*W tokens: GroupNorm(32, eps 1e-6) ->
                     to_q/to_k/to_v (bias) -> SDPA -> to_out (bias) -> + residual), ResnetBlock2D
    UpDecoderBlock2D(i) = 3 x ResnetBlock2D (first one changes the width) [+ Upsample2D: nearest x2 -> conv 3x3]
                     widths reversed(block_out_channels); no upsampler in the last block
    ResnetBlock2D (no time embedding) = GN(32, eps 1e-6) -> SiLU -> conv3x3 -> GN -> SiLU -> conv3x3, + shortcut
                     (1x1 ``conv_shortcut`` iff the width changes)
Sub-module names reproduce diffusers' state-dict keys (``post_quant_conv.weight``, ``decoder.mid_block.attentions.0.
to_q.weight``, ``decoder.up_blocks.2.resnets.0.conv_shortcut.weight`` ...).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class OracleVaeConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025          # sdxl-vae config.json


SDXL_VAE = OracleVaeConfig()
TINY_VAE = OracleVaeConfig(block_out_channels=(64, 64, 128, 128))


class _Resnet(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _Attention(nn.Module):      # diffusers Attention(heads=1, dim_head=C, bias=True, residual_connection=True)
    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        hs = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)          # (B, HW, C)
        q, k, v = self.to_q(hs), self.to_k(hs), self.to_v(hs)
        p = torch.softmax(q @ k.transpose(1, 2) / c ** 0.5, dim=-1)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class _Mid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(c, c, groups), _Resnet(c, c, groups)])
        self.attentions = nn.ModuleList([_Attention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _ConvHolder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)


class _Up(nn.Module):
    def __init__(self, cin, cout, n, groups, upsample):
        super().__init__()
        self.resnets = nn.ModuleList(_Resnet(cin if j == 0 else cout, cout, groups) for j in range(n))
        if upsample:
            self.upsamplers = nn.ModuleList([_ConvHolder(cout)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if hasattr(self, "upsamplers"):
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class _Decoder(nn.Module):
    def __init__(self, cfg: OracleVaeConfig):
        super().__init__()
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[-1], 3, padding=1)
        self.mid_block = _Mid(ch[-1], g)
        rev = list(reversed(ch))
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, c in enumerate(rev):
            self.up_blocks.append(_Up(prev, c, cfg.layers_per_block + 1, g, i < len(rev) - 1))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for up in self.up_blocks:
            x = up(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class OracleVaeDecoder(nn.Module):
    def __init__(self, cfg: OracleVaeConfig = SDXL_VAE):
        super().__init__()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = _Decoder(cfg)

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(self.post_quant_conv(z))

    @torch.no_grad()
    def forward(self, latents: torch.Tensor) -> torch.Tensor:
        """pipeline_diffsensei.py:359-363 with output_type "pt": latents / scaling_factor -> decode ->
        VaeImageProcessor.postprocess (denormalize: (x / 2 + 0.5).clamp(0, 1)).  (B, 3, 8h, 8w) in [0, 1]."""
        img = self.decode(latents / self.cfg.scaling_factor)
        return (img / 2 + 0.5).clamp(0, 1)
