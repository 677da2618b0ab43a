"""DDIM schedule (host side).  The reference uses whatever scheduler the checkpoint ships
(src/pipelines/pipeline_diffsensei.py:50,248-249,317,337); BASELINE.json fixes DDIM, so this mirrors diffusers'
``DDIMScheduler`` under the SDXL scheduler config: scaled_linear betas 0.00085..0.012 over 1000 train steps,
epsilon prediction, ``timestep_spacing="leading"``, ``steps_offset=1``, ``set_alpha_to_one=False``, no clipping,
eta = 0.  ``scale_model_input`` is the identity and ``init_noise_sigma`` is 1 for DDIM.  The per-step update
itself runs on the GPU, fused with the CFG blend (``ds_cfg_ddim_step``); this class only produces the
timesteps and the (alpha_prod_t, alpha_prod_t_prev) table the kernel reads.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


class DDIMScheduler:
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.timesteps: List[int] = []
        self.num_inference_steps = 0

    def set_timesteps(self, num_inference_steps: int, device=None) -> List[int]:
        ratio = self.num_train_timesteps // num_inference_steps
        self.num_inference_steps = num_inference_steps
        self.timesteps = [int(round(i * ratio)) + self.steps_offset for i in reversed(range(num_inference_steps))]
        return self.timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, t: int) -> Tuple[float, float]:
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def coefficient_table(self, device) -> torch.Tensor:
        """fp32 [T, 2] device tensor of (alpha_prod_t, alpha_prod_t_prev) in loop order."""
        return torch.tensor([self.coefficients(t) for t in self.timesteps], dtype=torch.float32, device=device)
