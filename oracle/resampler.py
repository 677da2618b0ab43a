"""Oracle restatement of the character Resampler (test infrastructure only).

Follows ``src/models/resampler.py`` of jianzongwu/DiffSensei: FeedForward (:11-18), PerceiverAttention
(:32-76), Resampler (:79-144).  Parameter names match the reference's ``state_dict`` exactly
(``latents``, ``proj_in``, ``proj_in_magi``, ``proj_out``, ``norm_out``, ``layers.{i}.0.{norm1,norm2,
to_q,to_kv,to_out}``, ``layers.{i}.1.{0,1,3}``, ``dummy_tokens``) so one set of weights drives the
reference, this oracle and the engine.  Pinned against the executed reference by
tests/golden/resampler_*.pt.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Perceiver(nn.Module):
    def __init__(self, dim: int, dim_head: int, heads: int):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.dim_head = heads, dim_head
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, 2 * inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, feats: torch.Tensor, lat: torch.Tensor) -> torch.Tensor:
        feats, lat = self.norm1(feats), self.norm2(lat)
        b, nq, _ = lat.shape
        h, d = self.heads, self.dim_head
        q = self.to_q(lat).view(b, nq, h, d).transpose(1, 2)
        kv = self.to_kv(torch.cat([feats, lat], dim=1))               # keys = image tokens then the latents
        k, v = (t.view(b, -1, h, d).transpose(1, 2) for t in kv.chunk(2, dim=-1))
        s = d ** -0.25                                                 # applied to q and k separately (:69-70)
        w = torch.softmax(((q * s) @ (k * s).transpose(-1, -2)).float(), dim=-1).to(q.dtype)
        return self.to_out((w @ v).transpose(1, 2).reshape(b, nq, h * d))


def _feed_forward(dim: int, mult: int) -> nn.Sequential:
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, dim * mult, bias=False), nn.GELU(),
                         nn.Linear(dim * mult, dim, bias=False))


class OracleResampler(nn.Module):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=4, num_dummy_tokens=4,
                 embedding_dim=768, magi_embedding_dim=512, output_dim=1024, ff_mult=4):
        super().__init__()
        self.num_queries, self.output_dim = num_queries, output_dim
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_in_magi = nn.Linear(magi_embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.layers = nn.ModuleList(
            nn.ModuleList([_Perceiver(dim, dim_head, heads), _feed_forward(dim, ff_mult)]) for _ in range(depth))
        self.dummy_tokens = nn.Parameter(torch.randn(num_dummy_tokens, output_dim))

    def forward(self, x: torch.Tensor, magi: torch.Tensor) -> torch.Tensor:
        bsz, n_ips, seq, _ = x.shape
        feats = self.proj_in(x.reshape(bsz * n_ips, seq, -1))
        feats = torch.cat([feats, self.proj_in_magi(magi).reshape(bsz * n_ips, 1, -1)], dim=1)
        lat = self.latents.expand(bsz * n_ips, -1, -1)
        for attn, ff in self.layers:
            lat = attn(feats, lat) + lat
            lat = ff(lat) + lat
        lat = self.norm_out(self.proj_out(lat)).reshape(bsz, n_ips * self.num_queries, self.output_dim)
        return torch.cat([self.dummy_tokens.unsqueeze(0).expand(bsz, -1, -1), lat], dim=1)
