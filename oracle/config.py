"""Topology description of the oracle UNet — TEST INFRASTRUCTURE ONLY, and deliberately INDEPENDENT of the
product package: nothing under ``oracle/`` imports ``diffsensei_b200`` (so ``import oracle`` never maps
``libdsengine.so``, and a wrong depth / width in the engine's own ``UNetConfig`` cannot be mirrored here by
construction).  The numbers are the published SDXL-base ``unet/config.json`` values the DiffSensei checkpoint
uses (block_out_channels 320/640/1280, transformer_layers_per_block 1/2/10 with the first level attention-free
``DownBlock2D`` -> 0/2/10, layers_per_block 2, attention_head_dim 5/10/20 = C/64, cross_attention_dim 2048,
norm_num_groups 32, addition_time_embed_dim 256, projection_class_embeddings_input_dim 2816) plus the three keys
``set_manga_modules`` registers (src/models/unet.py:50-53; configs/model/diffsensei.yaml: max_num_ips 4,
num_vision_tokens 16, max_num_dialogs 8).  ``tests/test_host_logic.py`` checks the engine's config against this
one field by field.
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Tuple


@dataclass(frozen=True)
class OracleUNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    transformer_layers_per_block: Tuple[int, ...] = (0, 2, 10)
    layers_per_block: int = 2
    attention_head_dim: int = 64
    cross_attention_dim: int = 2048
    norm_num_groups: int = 32
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    pooled_text_dim: int = 1280
    max_num_ips: int = 4
    num_vision_tokens: int = 16
    max_num_dialogs: int = 8

    @property
    def time_embed_dim(self) -> int:           # diffusers: block_out_channels[0] * 4
        return self.block_out_channels[0] * 4

    @property
    def num_ip_tokens(self) -> int:            # src/models/unet.py:79
        return self.max_num_ips * self.num_vision_tokens

    @property
    def num_dummy_tokens(self) -> int:         # src/models/unet.py:80
        return self.num_vision_tokens

    def heads(self, channels: int) -> int:
        return channels // self.attention_head_dim

    @classmethod
    def from_any(cls, cfg) -> "OracleUNetConfig":
        """Copy the same-named fields of any config object (e.g. the engine's UNetConfig in a parity test)."""
        if isinstance(cfg, cls):
            return cfg
        return cls(**{f.name: (tuple(getattr(cfg, f.name)) if isinstance(getattr(cfg, f.name), (list, tuple))
                               else getattr(cfg, f.name)) for f in fields(cls)})


SDXL = OracleUNetConfig()

# Same topology shrunk so the CPU oracle runs in well under a second (parity tests, smoke()).
TINY = OracleUNetConfig(block_out_channels=(64, 128, 256), transformer_layers_per_block=(0, 1, 2),
                        cross_attention_dim=128, projection_class_embeddings_input_dim=6 * 64 + 96,
                        addition_time_embed_dim=64, pooled_text_dim=96)


def feature_sizes(cfg, h: int, w: int):
    """(H, W) of the feature map at every resolution level: stride-2 / pad-1 3x3 convs => ceil halving."""
    res = [(h, w)]
    for _ in range(len(cfg.block_out_channels) - 1):
        res.append(((res[-1][0] - 1) // 2 + 1, (res[-1][1] - 1) // 2 + 1))
    return res


def unet_flops(cfg, B: int, h: int, w: int, hoist_kv: bool = False) -> float:
    """Analytic 2*MAC count of one UNetMangaModel.forward (conv 2*9*Cin*Cout*H*W*B, linear 2*in*out*tokens,
    SDPA 4*N*Nk*C*B) — the formula behind SURVEY.md §8d's 54.8 TFLOP for cfg2.  Walks the same topology the
    oracle module builds (down blocks, mid block, up blocks with skip concatenation)."""
    ch, depth = cfg.block_out_channels, cfg.transformer_layers_per_block
    n, td, L = len(ch), cfg.time_embed_dim, cfg.layers_per_block
    res = feature_sizes(cfg, h, w)
    n_cond = 77 + cfg.num_ip_tokens + cfg.num_dummy_tokens

    def resnet(cin, cout, lvl):
        px = res[lvl][0] * res[lvl][1] * B
        f = 2.0 * 9 * cin * cout * px + 2.0 * 9 * cout * cout * px + 2.0 * td * cout * B
        return f + (2.0 * cin * cout * px if cin != cout else 0.0)

    def transformer(c, d, lvl):
        N = res[lvl][0] * res[lvl][1]
        tok = N * B
        per = 4 * 2.0 * c * c * tok + 4.0 * N * N * c * B                     # attn1 q,k,v,out + SDPA
        per += 2 * 2.0 * c * c * tok + 4.0 * N * n_cond * c * B               # attn2 q,out + text & IP SDPA
        if not hoist_kv:
            per += 2 * 2.0 * cfg.cross_attention_dim * c * n_cond * B         # to_k/to_v + to_k_ip/to_v_ip
        per += 2.0 * c * 8 * c * tok + 2.0 * 4 * c * c * tok                  # GEGLU FF
        return 2 * 2.0 * c * c * tok + d * per                                # + proj_in / proj_out

    f = 2.0 * 9 * cfg.in_channels * ch[0] * h * w * B                         # conv_in
    prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(L):
            f += resnet(prev if j == 0 else c, c, i)
            if depth[i] > 0:
                f += transformer(c, depth[i], i)
        if i < n - 1:
            f += 2.0 * 9 * c * c * res[i + 1][0] * res[i + 1][1] * B          # Downsample2D conv (stride 2)
        prev = c
    f += 2 * resnet(ch[-1], ch[-1], n - 1) + transformer(ch[-1], depth[-1], n - 1)
    rev, rdepth = list(reversed(ch)), list(reversed(depth))
    prev = rev[0]
    for i, c in enumerate(rev):
        lvl = n - 1 - i
        skip_in = rev[min(i + 1, n - 1)]
        for j in range(L + 1):
            f += resnet((prev if j == 0 else c) + (skip_in if j == L else c), c, lvl)
            if rdepth[i] > 0:
                f += transformer(c, rdepth[i], lvl)
        if i < n - 1:
            f += 2.0 * 9 * c * c * res[lvl - 1][0] * res[lvl - 1][1] * B      # Upsample2D conv at the doubled size
        prev = c
    f += 2.0 * 9 * ch[0] * cfg.out_channels * h * w * B                       # conv_out
    f += 2.0 * B * (ch[0] * td + td * td + cfg.projection_class_embeddings_input_dim * td + td * td)
    return f
