"""Oracle DDIM schedule and denoise loop (test infrastructure only).

The reference never names its scheduler (it uses whatever the checkpoint ships,
src/pipelines/pipeline_diffsensei.py:50,248-249,317,337); BASELINE.json fixes DDIM.  This restates
diffusers' DDIMScheduler under the SDXL scheduler config — scaled_linear betas 0.00085..0.012 over 1000
train steps, epsilon prediction, timestep_spacing="leading", steps_offset=1, set_alpha_to_one=False,
no sample clipping, eta=0 (``scale_model_input`` is the identity, ``init_noise_sigma`` = 1) — and the loop
body of pipeline_diffsensei.py:310-337.  **Parity unpinned** (diffusers is not installed here).
"""
from __future__ import annotations

import torch


class DDIMSchedule:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]        # set_alpha_to_one=False
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset

    def set_timesteps(self, n: int):
        ratio = self.num_train_timesteps // n
        self.num_inference_steps = n
        self.timesteps = [int(round(i * ratio)) + self.steps_offset for i in reversed(range(n))]
        return self.timesteps

    def coefficients(self, t: int):
        """(alpha_prod_t, alpha_prod_t_prev) used by step()."""
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return float(a_t), float(a_prev)

    def step(self, eps: torch.Tensor, t: int, x: torch.Tensor) -> torch.Tensor:
        a_t, a_prev = self.coefficients(t)
        x0 = (x - (1.0 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_prev ** 0.5 * x0 + (1.0 - a_prev) ** 0.5 * eps


@torch.no_grad()
def denoise_loop(unet, latents, prompt_embeds, text_embeds, time_ids, bbox, aspect_ratio, dialog_bbox, guidance,
                 num_steps, schedule: DDIMSchedule | None = None, on_step=None):
    """pipeline_diffsensei.py:310-337 with CFG: conditions are already [negative ; positive] along batch."""
    schedule = schedule or DDIMSchedule()
    for i, t in enumerate(schedule.set_timesteps(num_steps)):
        model_in = torch.cat([latents] * 2)                                         # :315 (scale_model_input = id)
        eps = unet(model_in, t, prompt_embeds, text_embeds, time_ids, bbox, aspect_ratio, dialog_bbox)   # :322-329
        e_uncond, e_text = eps.chunk(2)                                             # :333
        eps = e_uncond + guidance * (e_text - e_uncond)                             # :334
        latents = schedule.step(eps, t, latents)                                    # :337
        if on_step is not None:
            on_step(i, t, latents)
    return latents
