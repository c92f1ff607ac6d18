"""Host side of the fused CFG + scheduler kernel (rg_sched_step): timestep lists and per-step coefficients.

Restates the two schedulers the reference ships:
  * PNDMScheduler with ``skip_prk_steps`` (PLMS) -- ``outputs/models/denoising/best/scheduler/scheduler_config.json``
    (denoise, colorize, sr), including the img2img quirk that ``counter`` restarts at 0 on a sliced timestep list;
  * DDIMScheduler, eta = 0 -- ``outputs/models/inpainting/best/scheduler/scheduler_config.json``.
The alpha-bar table is built with the same float32 torch ops diffusers uses (host-side scalar plumbing); the
per-element update runs in the CUDA kernel, with the 4-deep eps history and ``cur_sample`` kept in fp32 on the
device.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch


def alphas_cumprod(beta_start: float = 0.00085, beta_end: float = 0.012, n: int = 1000) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2     # scaled_linear
    return torch.cumprod(1.0 - betas, dim=0)


@dataclass
class StepPlan:
    """One launch of rg_sched_step."""
    timestep: int                    # value fed to the UNet
    store_slot: int                  # history slot that receives the guided eps (-1: none)
    w: list = field(default_factory=lambda: [0.0] * 5)     # weights of slots 0..3 and of the current eps
    use_cur: bool = False
    save_cur: bool = False
    c_sample: float = 1.0
    c_eps: float = 0.0


class _SchedulerBase:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1, set_alpha_to_one: bool = False, **_ignored):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps: list[int] = []
        self.num_inference_steps = 0

    @classmethod
    def from_config(cls, cfg: dict):
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    def _alpha(self, t: int) -> torch.Tensor:
        return self.alphas_cumprod[t] if t >= 0 else self.final_alpha_cumprod

    def add_noise_coeffs(self, t: int) -> tuple[float, float]:
        a = self.alphas_cumprod[t]
        return float(a ** 0.5), float((1 - a) ** 0.5)

    def get_timesteps(self, num_inference_steps: int, strength: float) -> list[int]:
        """Img2Img / Inpaint ``get_timesteps``: drop the first ``N - int(N * strength)`` entries."""
        init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
        t_start = max(num_inference_steps - init_timestep, 0)
        return self.timesteps[t_start * self.order:]


class PNDMScheduler(_SchedulerBase):
    kind = "pndm"

    def set_timesteps(self, num_inference_steps: int):
        self.num_inference_steps = num_inference_steps
        r = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * r).round() + self.steps_offset
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1]
        self.timesteps = [int(t) for t in plms]

    def _coeffs(self, t: int, prev_t: int) -> tuple[float, float]:
        a_t, a_p = self._alpha(t), self._alpha(prev_t)
        b_t, b_p = 1 - a_t, 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return float(sample_coeff), float((a_p - a_t) / denom)

    def plan(self, timesteps: list[int]) -> list[StepPlan]:
        """Unroll ``step_plms`` over the (possibly sliced) timestep list into kernel launches."""
        r = self.num_train_timesteps // self.num_inference_steps
        ets: list[int] = []           # history as slot indices, oldest first
        n_stored = 0
        plans = []
        for counter, t_in in enumerate(timesteps):
            t, prev_t = t_in, t_in - r
            p = StepPlan(timestep=t_in, store_slot=-1)
            if counter != 1:
                ets = ets[-3:]
                slot = n_stored % 4
                n_stored += 1
                ets.append(slot)
                p.store_slot = slot
            else:
                prev_t, t = t, t + r
            if len(ets) == 1 and counter == 0:
                p.w[4] = 1.0
                p.save_cur = True
            elif len(ets) == 1 and counter == 1:
                p.w[4] = 0.5
                p.w[ets[-1]] = 0.5
                p.use_cur = True
            elif len(ets) == 2:
                p.w[4], p.w[ets[-2]] = 3 / 2, -1 / 2
            elif len(ets) == 3:
                p.w[4], p.w[ets[-2]], p.w[ets[-3]] = 23 / 12, -16 / 12, 5 / 12
            else:
                p.w[4], p.w[ets[-2]], p.w[ets[-3]], p.w[ets[-4]] = 55 / 24, -59 / 24, 37 / 24, -9 / 24
            p.c_sample, p.c_eps = self._coeffs(t, prev_t)
            plans.append(p)
        return plans


class DDIMScheduler(_SchedulerBase):
    kind = "ddim"

    def set_timesteps(self, num_inference_steps: int):
        self.num_inference_steps = num_inference_steps
        r = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * r).round()[::-1].astype(np.int64) + self.steps_offset
        self.timesteps = [int(t) for t in ts]

    def plan(self, timesteps: list[int]) -> list[StepPlan]:
        r = self.num_train_timesteps // self.num_inference_steps
        plans = []
        for t in timesteps:
            a_t, a_p = float(self._alpha(t)), float(self._alpha(t - r))
            # prev = sqrt(a_p) * (x - sqrt(1-a_t) e) / sqrt(a_t) + sqrt(1-a_p) e          (eta = 0)
            c_sample = (a_p / a_t) ** 0.5
            c_eps = (a_p ** 0.5) * ((1 - a_t) ** 0.5) / (a_t ** 0.5) - (1 - a_p) ** 0.5
            p = StepPlan(timestep=t, store_slot=-1, c_sample=c_sample, c_eps=c_eps)
            p.w[4] = 1.0
            plans.append(p)
        return plans


SCHEDULERS = {"PNDMScheduler": PNDMScheduler, "DDIMScheduler": DDIMScheduler}
