"""Oracle restatement of UNetMangaModel (test infrastructure only; CPU fp32 torch).

Follows the reference's ``src/models/unet.py``:
  * module install / IP weights cloned from to_k/to_v / ``dialog_bbox_embedding``  <- set_manga_modules (:44-86)
  * ``encode_dialog_bbox``                                                         <- :88-114
  * forward order                                                                  <- :186-338
and, for the blocks that live in the un-vendored, un-pinned ``diffusers`` dependency,
the published SDXL semantics: ResnetBlock2D, Transformer2DModel (linear projections),
BasicTransformerBlock, GEGLU FeedForward, Timesteps / TimestepEmbedding, Down/Upsample2D.
**Parity unpinned** for those diffusers blocks: no reference output exists in this environment.

Sub-module names reproduce diffusers' state-dict keys (``down_blocks.1.attentions.0.transformer_blocks.0.
attn2.to_q.weight`` ...) and the reference's additions (``...attn2.processor.to_k_ip.weight``,
``dialog_bbox_embedding``; src/models/unet.py:73-86) so one state dict drives oracle and engine.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import attention as A
from .config import OracleUNetConfig


def timestep_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin](t * w_i)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.to(torch.float32)[:, None] * freqs[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class TimestepMLP(nn.Module):      # diffusers TimestepEmbedding
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class Resnet(nn.Module):           # diffusers ResnetBlock2D (eps 1e-5, swish, output_scale_factor 1)
    def __init__(self, cin, cout, temb_dim, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class AttnShell(nn.Module):        # diffusers Attention: to_q/k/v without bias, to_out = [Linear(+bias), Dropout]
    def __init__(self, dim, kv_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(kv_dim, dim, bias=False)
        self.to_v = nn.Linear(kv_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])


class IPProcessorParams(nn.Module):   # parameters MaskedIPAttnProcessor2_0 owns (attention_processor.py:112-113)
    def __init__(self, dim, kv_dim):
        super().__init__()
        self.to_k_ip = nn.Linear(kv_dim, dim, bias=False)
        self.to_v_ip = nn.Linear(kv_dim, dim, bias=False)
        self.scale = 1.0


class GEGLUProj(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Linear(dim, dim * 8)


class FeedForward(nn.Module):      # diffusers FeedForward(activation_fn="geglu"): net = [GEGLU, Dropout, Linear]
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLUProj(dim), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        val, gate = self.net[0].proj(x).chunk(2, dim=-1)
        return self.net[2](val * F.gelu(gate))            # exact (erf) GELU


class TransformerBlock(nn.Module):  # diffusers BasicTransformerBlock
    def __init__(self, dim, heads, kv_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = AttnShell(dim, dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = AttnShell(dim, kv_dim, heads)
        self.attn2.processor = IPProcessorParams(dim, kv_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, hs, ehs, bbox, aspect_ratio, cfg: OracleUNetConfig):
        a1, a2 = self.attn1, self.attn2
        hs = hs + A.self_attention(self.norm1(hs), a1.to_q.weight, a1.to_k.weight, a1.to_v.weight,
                                   a1.to_out[0].weight, a1.to_out[0].bias, a1.heads)
        pr = a2.processor
        hs = hs + A.cross_ip_attention(self.norm2(hs), ehs, bbox, aspect_ratio, a2.to_q.weight, a2.to_k.weight,
                                       a2.to_v.weight, pr.to_k_ip.weight, pr.to_v_ip.weight, a2.to_out[0].weight,
                                       a2.to_out[0].bias, a2.heads, pr.scale, cfg.num_ip_tokens, cfg.num_dummy_tokens)
        return hs + self.ff(self.norm3(hs))


class Transformer2D(nn.Module):     # diffusers Transformer2DModel, use_linear_projection=True
    def __init__(self, dim, heads, depth, kv_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList(TransformerBlock(dim, heads, kv_dim) for _ in range(depth))
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ehs, bbox, aspect_ratio, cfg):
        b, c, h, w = x.shape
        hs = self.proj_in(self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c))
        for blk in self.transformer_blocks:
            hs = blk(hs, ehs, bbox, aspect_ratio, cfg)
        return self.proj_out(hs).reshape(b, h, w, c).permute(0, 3, 1, 2) + x


class ConvHolder(nn.Module):        # Downsample2D / Upsample2D keep their conv under ".conv"
    def __init__(self, c, stride):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1)


class Stage(nn.Module):
    """One Down/Up block: resnets (+ attentions) (+ down/upsampler)."""

    def __init__(self, resnet_io, temb_dim, groups, depth, heads, kv_dim, sampler: str | None):
        super().__init__()
        self.resnets = nn.ModuleList(Resnet(i, o, temb_dim, groups) for i, o in resnet_io)
        cout = resnet_io[-1][1]
        if depth > 0:
            self.attentions = nn.ModuleList(Transformer2D(cout, heads, depth, kv_dim, groups) for _ in resnet_io)
        else:
            self.attentions = None
        if sampler == "down":
            self.downsamplers = nn.ModuleList([ConvHolder(cout, 2)])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([ConvHolder(cout, 1)])


class MidStage(nn.Module):          # UNetMidBlock2DCrossAttn: resnet, attn, resnet
    def __init__(self, c, temb_dim, groups, depth, heads, kv_dim):
        super().__init__()
        self.resnets = nn.ModuleList([Resnet(c, c, temb_dim, groups), Resnet(c, c, temb_dim, groups)])
        self.attentions = nn.ModuleList([Transformer2D(c, heads, depth, kv_dim, groups)])


def dialog_boxes_to_pixels(dialog_bbox: torch.Tensor, height: int, width: int):
    """int() truncation of bbox*size in the tensor's own dtype, clamp to the image (unet.py:102-108)."""
    boxes = []
    for j in range(dialog_bbox.shape[0]):
        x1 = int(dialog_bbox[j, 0] * width)
        y1 = int(dialog_bbox[j, 1] * height)
        x2 = int(dialog_bbox[j, 2] * width)
        y2 = int(dialog_bbox[j, 3] * height)
        boxes.append((max(0, x1), max(0, y1), min(width, x2), min(height, y2)))
    return boxes


def encode_dialog_bbox(sample: torch.Tensor, dialog_bbox: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """sample + emb * 1[pixel in union of half-open boxes]  (unet.py:88-114)."""
    b, c, h, w = sample.shape
    add = torch.zeros_like(sample)
    for i in range(b):
        for (x1, y1, x2, y2) in dialog_boxes_to_pixels(dialog_bbox[i], h, w):
            add[i, :, y1:y2, x1:x2] = emb.view(c, 1, 1)
    return sample + add


class OracleUNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        cfg = OracleUNetConfig.from_any(cfg)          # own dataclass: the oracle never imports the product package
        self.cfg = cfg
        ch, g, kv = cfg.block_out_channels, cfg.norm_num_groups, cfg.cross_attention_dim
        td, depth = cfg.time_embed_dim, cfg.transformer_layers_per_block
        n = len(ch)
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepMLP(ch[0], td)
        self.add_embedding = TimestepMLP(cfg.projection_class_embeddings_input_dim, td)
        self.down_blocks = nn.ModuleList()
        prev = ch[0]
        for i, c in enumerate(ch):
            io = [(prev if j == 0 else c, c) for j in range(cfg.layers_per_block)]
            self.down_blocks.append(Stage(io, td, g, depth[i], cfg.heads(c), kv, "down" if i < n - 1 else None))
            prev = c
        self.mid_block = MidStage(ch[-1], td, g, depth[-1], cfg.heads(ch[-1]), kv)
        rev, rdepth = list(reversed(ch)), list(reversed(depth))
        self.up_blocks = nn.ModuleList()
        prev = rev[0]
        for i, c in enumerate(rev):
            skip_in = rev[min(i + 1, n - 1)]
            nl = cfg.layers_per_block + 1
            io = [((prev if j == 0 else c) + (skip_in if j == nl - 1 else c), c) for j in range(nl)]
            self.up_blocks.append(Stage(io, td, g, rdepth[i], cfg.heads(c), kv, "up" if i < n - 1 else None))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)
        # set_manga_modules (unet.py:72-86): IP projections start as clones of to_k/to_v; dialog embedding ~ randn
        for m in self.modules():
            if isinstance(m, TransformerBlock):
                m.attn2.processor.to_k_ip.weight.data.copy_(m.attn2.to_k.weight.data)
                m.attn2.processor.to_v_ip.weight.data.copy_(m.attn2.to_v.weight.data)
        self.dialog_bbox_embedding = nn.Parameter(torch.randn(ch[0]))

    def set_ip_scale(self, scale: float):       # pipeline_diffsensei.py:172-178
        for m in self.modules():
            if isinstance(m, IPProcessorParams):
                m.scale = scale

    def embed_time(self, timestep, text_embeds, time_ids, batch):
        cfg = self.cfg
        w = self.conv_in.weight                      # the oracle runs wherever (and in whatever dtype) its weights live
        t = torch.as_tensor(timestep, dtype=torch.float32, device=w.device).reshape(-1).expand(batch)
        emb = self.time_embedding(timestep_sinusoid(t, cfg.block_out_channels[0]).to(w.dtype))
        tid = timestep_sinusoid(time_ids.reshape(-1), cfg.addition_time_embed_dim).reshape(batch, -1).to(w.dtype)
        return emb + self.add_embedding(torch.cat([text_embeds, tid], dim=-1))     # unet.py:190-196

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, text_embeds, time_ids, bbox, aspect_ratio,
                dialog_bbox=None):
        cfg = self.cfg
        nlev = len(cfg.block_out_channels)
        emb = self.embed_time(timestep, text_embeds, time_ids, sample.shape[0])
        need_size = any(d % (2 ** (nlev - 1)) != 0 for d in sample.shape[-2:])   # unet.py:152-162
        x = self.conv_in(sample)
        if dialog_bbox is not None:
            x = encode_dialog_bbox(x, dialog_bbox, self.dialog_bbox_embedding)     # unet.py:208-210
        skips = [x]
        ehs = encoder_hidden_states
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                x = res(x, emb)
                if blk.attentions is not None:
                    x = blk.attentions[j](x, ehs, bbox, aspect_ratio, cfg)
                skips.append(x)
            if hasattr(blk, "downsamplers"):
                x = blk.downsamplers[0].conv(x)
                skips.append(x)
        x = self.mid_block.resnets[0](x, emb)
        x = self.mid_block.attentions[0](x, ehs, bbox, aspect_ratio, cfg)
        x = self.mid_block.resnets[1](x, emb)
        for i, blk in enumerate(self.up_blocks):
            nres = len(blk.resnets)
            mine, skips = skips[-nres:], skips[:-nres]
            for j, res in enumerate(blk.resnets):
                x = res(torch.cat([x, mine[-1 - j]], dim=1), emb)
                if blk.attentions is not None:
                    x = blk.attentions[j](x, ehs, bbox, aspect_ratio, cfg)
            if hasattr(blk, "upsamplers"):
                if need_size:                                                      # unet.py:312-313
                    x = F.interpolate(x, size=skips[-1].shape[2:], mode="nearest")
                else:
                    x = F.interpolate(x, scale_factor=2.0, mode="nearest")
                x = blk.upsamplers[0].conv(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))
