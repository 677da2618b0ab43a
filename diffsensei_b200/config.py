"""Static description of the denoiser topology (the subset of diffusers' UNet2DConditionModel config that
the DiffSensei checkpoint uses, plus the three keys ``set_manga_modules`` registers,
src/models/unet.py:50-53).  Pure data: shared by the engine, the weight factory and the test oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    # transformer depth per resolution level; 0 = plain Down/UpBlock2D without attention
    transformer_layers_per_block: Tuple[int, ...] = (0, 2, 10)
    layers_per_block: int = 2
    attention_head_dim: int = 64          # per-head width (SDXL: heads = C / 64 = 5, 10, 20)
    cross_attention_dim: int = 2048
    norm_num_groups: int = 32
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816   # 6*256 + 1280 (text_time)
    pooled_text_dim: int = 1280
    # manga modules (configs/model/diffsensei.yaml)
    max_num_ips: int = 4
    num_vision_tokens: int = 16
    max_num_dialogs: int = 8

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def num_ip_tokens(self) -> int:        # MaskedIPAttnProcessor2_0.num_ip_tokens (unet.py:79)
        return self.max_num_ips * self.num_vision_tokens

    @property
    def num_dummy_tokens(self) -> int:     # unet.py:80
        return self.num_vision_tokens

    def heads(self, channels: int) -> int:
        return channels // self.attention_head_dim


SDXL_MANGA = UNetConfig()

# Same topology, shrunk so that the CPU oracle runs in well under a second: used by parity tests.
TINY = UNetConfig(block_out_channels=(64, 128, 256), transformer_layers_per_block=(0, 1, 2),
                  cross_attention_dim=128, projection_class_embeddings_input_dim=6 * 64 + 96,
                  addition_time_embed_dim=64, pooled_text_dim=96)


@dataclass(frozen=True)
class ResamplerConfig:
    """configs/model/diffsensei.yaml + scripts/demo/gradio_wo_mllm.py:174-185."""
    dim: int = 1280
    depth: int = 4
    dim_head: int = 64
    heads: int = 20
    num_queries: int = 16
    num_dummy_tokens: int = 16
    embedding_dim: int = 1280        # CLIP ViT-H hidden size
    magi_embedding_dim: int = 768    # Magi ViT-MAE hidden size
    output_dim: int = 2048
    ff_mult: int = 4


RESAMPLER = ResamplerConfig()
RESAMPLER_TINY = ResamplerConfig(dim=128, depth=2, heads=2, num_queries=16, num_dummy_tokens=16, embedding_dim=64,
                                 magi_embedding_dim=32, output_dim=128)


@dataclass(frozen=True)
class VaeConfig:
    """AutoencoderKL (sdxl-vae ``config.json``) — the decoder half the pipeline uses after the denoise loop
    (src/pipelines/pipeline_diffsensei.py:339-363)."""
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025
    force_upcast: bool = True


SDXL_VAE = VaeConfig()
TINY_VAE = VaeConfig(block_out_channels=(64, 64, 128, 128))
