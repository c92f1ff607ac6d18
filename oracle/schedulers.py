"""Restatement of diffusers' ``PNDMScheduler`` (PLMS mode) and ``DDIMScheduler`` (eta=0).

Test infrastructure only (see ``oracle/__init__.py``).

Configs: ``/root/reference/outputs/models/denoising/best/scheduler/scheduler_config.json:1-15``
(PNDM, ``skip_prk_steps`` true, ``steps_offset`` 1, leading spacing) for denoise /
colorize / sr and ``outputs/models/inpainting/best/scheduler/scheduler_config.json:1-20``
(DDIM) for inpaint.  Arithmetic is kept in torch float32 exactly as upstream does
(alpha-bar table is a float32 tensor; scalars are 0-dim float32 tensors).
"""
from __future__ import annotations

import numpy as np
import torch


def _alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class _Base:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 steps_offset=1, set_alpha_to_one=False):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = _alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original.device, dtype=original.dtype)
        timesteps = timesteps.to(original.device)
        sa = (ac[timesteps] ** 0.5).flatten()
        sb = ((1 - ac[timesteps]) ** 0.5).flatten()
        while sa.ndim < original.ndim:
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original + sb * noise


class PNDMScheduler(_Base):
    """PLMS: 4th-order linear multistep with the counter==1 Heun-style correction."""
    pndm_order = 4

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        _ts = (np.arange(0, num_inference_steps) * step_ratio).round() + self.steps_offset
        plms = np.concatenate([_ts[:-1], _ts[-2:-1], _ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def step(self, model_output, timestep, sample):
        timestep = int(timestep)
        r = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - r
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_timestep = timestep
            timestep = timestep + r
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2]
                                       + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev = self._get_prev_sample(sample, timestep, prev_timestep, model_output)
        self.counter += 1
        return prev

    def _get_prev_sample(self, sample, timestep, prev_timestep, model_output):
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff * sample - (a_p - a_t) * model_output / denom


class DDIMScheduler(_Base):
    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.steps_offset)

    def step(self, model_output, timestep, sample, eta: float = 0.0):
        timestep = int(timestep)
        prev_timestep = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        # eta == 0 -> std_dev_t == 0 (the reference never passes eta; pipelines default to 0.0)
        direction = (1 - a_p) ** 0.5 * model_output
        return a_p ** 0.5 * x0 + direction


def get_timesteps(scheduler, num_inference_steps: int, strength: float):
    """Img2Img / Inpaint pipelines' ``get_timesteps``."""
    init_timestep = min(int(num_inference_steps * strength), num_inference_steps)
    t_start = max(num_inference_steps - init_timestep, 0)
    timesteps = scheduler.timesteps[t_start * scheduler.order:]
    return timesteps, num_inference_steps - t_start
