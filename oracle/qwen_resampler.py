"""Oracle restatement of the MLLM adaptor's ``QwenResampler`` (test infrastructure only).

Follows src/models/qwen_resampler.py of jianzongwu/DiffSensei: the fixed 2-D sin-cos position table (:37-85, numpy
float32 maths as in the reference), ``grid_size**2`` learned queries, ``kv_proj`` / ``ln_kv`` / ``ln_q`` and ONE
``nn.MultiheadAttention`` (:87-145).  Same state-dict keys.  Pinned against the executed reference by
tests/golden/qwen_resampler.pt (tools/make_golden.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:          # get_1d_sincos_pos_embed_from_grid (:67-85)
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.
    omega = 1. / 10000 ** omega
    out = np.einsum('m,d->md', pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim: int, grid_size: int) -> np.ndarray:            # get_2d_sincos_pos_embed (:37-64): w first
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    return np.concatenate([sincos_1d(embed_dim // 2, grid[0]), sincos_1d(embed_dim // 2, grid[1])], axis=1)


class OracleQwenResampler(nn.Module):
    def __init__(self, grid_size, embed_dim, num_heads, kv_dim=None):
        super().__init__()
        self.num_queries = grid_size ** 2
        self.pos_embed = nn.Parameter(torch.from_numpy(sincos_2d(embed_dim, grid_size)).float(), requires_grad=False)
        self.query = nn.Parameter(torch.zeros(self.num_queries, embed_dim))
        self.kv_proj = nn.Linear(kv_dim, embed_dim, bias=False) if (kv_dim is not None and kv_dim != embed_dim) \
            else nn.Identity()
        self.attn = nn.MultiheadAttention(embed_dim, num_heads)
        self.ln_q = nn.LayerNorm(embed_dim)
        self.ln_kv = nn.LayerNorm(embed_dim)

    @torch.no_grad()
    def forward(self, x):
        if x.size(1) != self.num_queries:
            raise NotImplementedError("bicubic interpolation of the position table (get_abs_pos) is not restated")
        x = self.ln_kv(self.kv_proj(x)).permute(1, 0, 2)                 # (L, B, E)
        n = x.shape[1]
        q = self.ln_q(self.query).unsqueeze(1).repeat(1, n, 1) + self.pos_embed.unsqueeze(1)
        return self.attn(q, x + self.pos_embed.unsqueeze(1), x)[0].permute(1, 0, 2)


def seeded_case(kwargs: dict, seed: int):
    """The weights / input of tests/golden/qwen_resampler.pt, regenerated (tools/make_golden.py uses the same recipe in
    the reference module's state-dict order, which this module reproduces)."""
    m = OracleQwenResampler(**kwargs).eval()
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in m.state_dict().items():
        if k == "pos_embed":
            sd[k] = v
        elif k.endswith("weight") and v.dim() == 1:
            sd[k] = (1 + 0.1 * torch.randn(v.shape, generator=g)).to(torch.bfloat16).float()
        elif k.endswith("bias"):
            sd[k] = (0.05 * torch.randn(v.shape, generator=g)).to(torch.bfloat16).float()
        elif k == "query":
            sd[k] = torch.randn(v.shape, generator=g).to(torch.bfloat16).float()
        else:
            sd[k] = (torch.randn(v.shape, generator=g) * v.shape[-1] ** -0.5).to(torch.bfloat16).float()
    m.load_state_dict(sd)
    x = torch.randn(2, kwargs["grid_size"] ** 2, kwargs["kv_dim"], generator=g).to(torch.bfloat16).float()
    return m, sd, x
