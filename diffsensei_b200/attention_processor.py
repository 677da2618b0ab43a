"""Engine-backed attention processors with the reference's diffusers AttnProcessor protocol.

Same class names, constructor arguments, parameters and call signature as
``src/models/attention_processor.py`` of jianzongwu/DiffSensei:

    AttnProcessor2_0()                                                                  (:7-96)
    MaskedIPAttnProcessor2_0(hidden_size, cross_attention_dim, scale, num_ip_tokens, num_dummy_tokens)
        .to_k_ip / .to_v_ip : nn.Linear(cross_attention_dim, hidden_size, bias=False)   (:100-113)
        .scale              : mutable float, found via hasattr by pipeline.set_ip_scale (pipeline :172-178)
    proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, bbox=None,
         [dialog_bbox=None,] aspect_ratio=None, *args, **kwargs) -> Tensor(B, N, C)     (:19-31, :171-182)

so they can be installed on a stock diffusers ``UNet2DConditionModel`` with ``set_attn_processor`` (the
innermost drop-in seam) — ``load_ip_adapter``'s ``ModuleList(unet.attn_processors.values())``
state-dict indexing (src/models/utils.py:46-48) keeps working because both are ``nn.Module`` s with the same
parameter names.  The arithmetic runs in libdsengine: one fused-QKV tcgen05 GEMM + the flash kernel for
self-attention; for cross-attention the text / IP K|V projections (cached per conditioning tensor), ONE fused
kernel for both softmaxes, the bbox mask and the ``scale`` blend, and the output projection with its bias
(and optional residual) in the GEMM epilogue.  bf16, head_dim 64, 3-D hidden states — the SDXL configuration;
anything else raises (there is no eager fallback).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops

bf16, f32 = torch.bfloat16, torch.float32


def _unsupported(attn, hidden_states, attention_mask):
    if hidden_states.dim() != 3:
        raise NotImplementedError("engine processors take (B, N, C) hidden states (SDXL transformer blocks)")
    if attention_mask is not None:
        raise NotImplementedError("attention_mask is None on the DiffSensei sampling path (unet.py:172-183)")
    for name in ("spatial_norm", "group_norm"):
        if getattr(attn, name, None) is not None:
            raise NotImplementedError(f"attn.{name} is not used by the SDXL Attention shell")
    if getattr(attn, "norm_cross", False):
        raise NotImplementedError("attn.norm_cross is not used by the SDXL Attention shell")
    if hidden_states.dtype != bf16 or not hidden_states.is_cuda:
        raise ops.DsEngineError("engine processors need bf16 CUDA hidden states (no CPU / fp32 fallback)")


class _PackCache:
    """Fused/packed copies of the owning Attention module's weights, rebuilt when the parameters change."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, tensors, build):
        key = tuple((t.data_ptr(), t._version, t.dtype) for t in tensors)
        if key != self._key:
            self._val = build()
            self._key = key
        return self._val


def _out_proj(attn, a, residual):
    lin = attn.to_out[0]
    bias = None if lin.bias is None else lin.bias.detach().to(f32).contiguous()
    scale = 1.0 / float(getattr(attn, "rescale_output_factor", 1.0))
    res = residual.contiguous() if getattr(attn, "residual_connection", False) else None
    return ops.gemm(a, lin.weight.detach().to(bf16).contiguous(), bias, residual=res,
                    out_scale=0.0 if scale == 1.0 else scale)


class AttnProcessor2_0(nn.Module):
    def __init__(self):
        super().__init__()
        self._cache = _PackCache()

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, bbox=None,
                 dialog_bbox=None, aspect_ratio=None, *args, **kwargs):
        _unsupported(attn, hidden_states, attention_mask)
        if encoder_hidden_states is not None:
            raise NotImplementedError("AttnProcessor2_0 is installed on attn1 (self-attention) sites only "
                                      "(src/models/unet.py:68-69)")
        hs = hidden_states.contiguous()
        ws = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight)
        wqkv = self._cache.get(ws, lambda: torch.cat([w.detach() for w in ws], 0).to(bf16).contiguous())
        a = ops.attention_self(ops.gemm(hs, wqkv), attn.heads)
        return _out_proj(attn, a, hs)


class MaskedIPAttnProcessor2_0(nn.Module):
    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_ip_tokens=4, num_dummy_tokens=4):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        self.num_ip_tokens = num_ip_tokens
        self.num_dummy_tokens = num_dummy_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self._w_text = _PackCache()
        self._w_ip = _PackCache()
        self._kv = _PackCache()

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, bbox=None,
                 aspect_ratio=None, *args, **kwargs):
        _unsupported(attn, hidden_states, attention_mask)
        if encoder_hidden_states is None or bbox is None or aspect_ratio is None:
            raise ValueError("MaskedIPAttnProcessor2_0 needs encoder_hidden_states, bbox and aspect_ratio "
                             "(cross_attention_kwargs, src/pipelines/pipeline_diffsensei.py:270-273)")
        hs = hidden_states.contiguous()
        ehs = encoder_hidden_states
        end = ehs.shape[1] - (self.num_ip_tokens + self.num_dummy_tokens)        # reference :213
        wt = (attn.to_k.weight, attn.to_v.weight)
        wi = (self.to_k_ip.weight, self.to_v_ip.weight)
        w_text = self._w_text.get(wt, lambda: torch.cat([w.detach() for w in wt], 0).to(bf16).contiguous())
        w_ip = self._w_ip.get(wi, lambda: torch.cat([w.detach() for w in wi], 0).to(bf16).contiguous())

        def project():   # timestep-invariant: recomputed only when the conditioning tensor or weights change
            e = ehs.detach().to(bf16)
            return (ops.gemm(e[:, :end].contiguous(), w_text), ops.gemm(e[:, end:].contiguous(), w_ip))

        kv_text, kv_ip = self._kv.get((ehs,) + wt + wi, project)
        q = ops.gemm(hs, attn.to_q.weight.detach().to(bf16).contiguous())
        num_ips = bbox.shape[1]
        a = ops.attention_cross_ip(q, kv_text, kv_ip, bbox.detach().to(device=hs.device, dtype=f32).contiguous(),
                                   attn.heads, float(aspect_ratio), float(self.scale),
                                   self.num_ip_tokens // num_ips, self.num_dummy_tokens)
        return _out_proj(attn, a, hs)


class _SiteView:
    """What ``UNetMangaEngine.attn_processors`` hands out per site: API-shaped, not on the compute path."""

    def __init__(self, engine, is_cross: bool):
        self._engine = engine
        if is_cross:
            self.scale = engine.ip_scale


def build_processor_table(engine) -> Dict[str, object]:
    from .weights import transformer_sites
    table: Dict[str, object] = {}
    for p, _c, depth in transformer_sites(engine.cfg):
        for k in range(depth):
            table[f"{p}.transformer_blocks.{k}.attn1.processor"] = _SiteView(engine, False)
            table[f"{p}.transformer_blocks.{k}.attn2.processor"] = _SiteView(engine, True)
    return table
