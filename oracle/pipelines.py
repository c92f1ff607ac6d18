"""Restatement of diffusers' ``StableDiffusionImg2ImgPipeline.__call__`` and
``StableDiffusionInpaintPipeline.__call__`` (0.35.x) plus ``VaeImageProcessor``.

Test infrastructure only (see ``oracle/__init__.py``).

These are the calls the reference makes at ``/root/reference/src/inference.py:486-494``
(denoise), ``:566-573`` (sr), ``:664-672`` (colorize) and ``:758-767`` (inpaint).  The
control flow follows SURVEY.md Appendix A.1 / A.2 / A.6 step by step; every RNG draw
is made with the caller's ``torch.Generator`` in upstream order and shape.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

from .schedulers import PNDMScheduler, DDIMScheduler, get_timesteps
from .vae import randn_tensor

VAE_SCALE = 8


# ----------------------------------------------------------------------------- VaeImageProcessor
def preprocess_image(image: Image.Image, height: int | None = None, width: int | None = None) -> torch.Tensor:
    """``VaeImageProcessor.preprocess`` for one PIL image -> f32 [1,3,H,W] in [-1,1]."""
    if height is None:
        height = image.height
    if width is None:
        width = image.width
    width, height = (x - x % VAE_SCALE for x in (width, height))
    image = image.resize((width, height), resample=Image.LANCZOS)
    arr = np.array(image).astype(np.float32) / 255.0
    if arr.ndim == 2:
        arr = arr[..., None]
    t = torch.from_numpy(arr[None].transpose(0, 3, 1, 2).copy())
    return 2.0 * t - 1.0


def preprocess_mask(mask: Image.Image, height: int, width: int) -> torch.Tensor:
    """Mask processor: ``do_normalize=False, do_binarize=True, do_convert_grayscale=True``."""
    width, height = (x - x % VAE_SCALE for x in (width, height))
    mask = mask.resize((width, height), resample=Image.LANCZOS).convert("L")
    arr = np.array(mask).astype(np.float32) / 255.0
    t = torch.from_numpy(arr[None, None].copy())
    t[t < 0.5] = 0
    t[t >= 0.5] = 1
    return t


def postprocess_image(t: torch.Tensor) -> np.ndarray:
    """``postprocess(..., output_type='pil')`` up to the uint8 HWC array."""
    t = (t / 2 + 0.5).clamp(0, 1)
    arr = t.cpu().permute(0, 2, 3, 1).float().numpy()
    return (arr * 255).round().astype("uint8")


@dataclass
class Trace:
    """Per-step record used by the parity tests."""
    timesteps: list = field(default_factory=list)
    unet_in: list = field(default_factory=list)      # latents fed to the UNet at each step [B,4|9,h,w]
    eps: list = field(default_factory=list)          # guided eps [B,4,h,w]
    latents: list = field(default_factory=list)      # scheduler output
    init_latents: torch.Tensor | None = None
    final_latents: torch.Tensor | None = None
    decoded: torch.Tensor | None = None


class OraclePipeline:
    """unet / vae are the oracle modules; ``prompt_embeds`` / ``negative_prompt_embeds`` are
    ``[1,77,768]`` CLIP outputs (the text encoder itself stays in ``transformers``)."""

    def __init__(self, unet, vae, scheduler_kind: str):
        self.unet, self.vae = unet, vae
        self.scheduler_kind = scheduler_kind

    def _scheduler(self):
        return PNDMScheduler() if self.scheduler_kind == "pndm" else DDIMScheduler()

    def _loop(self, latents, timesteps, sched, embeds, g, extra_cond, trace):
        do_cfg = g > 1.0
        for t in timesteps:
            x = torch.cat([latents] * 2) if do_cfg else latents
            if extra_cond is not None:
                x = torch.cat([x, extra_cond[0], extra_cond[1]], dim=1)
            eps = self.unet(x, t, embeds)
            if do_cfg:
                u, c = eps.chunk(2)
                eps = u + g * (c - u)
            if trace is not None:
                trace.timesteps.append(int(t))
                trace.unet_in.append(x.clone())
                trace.eps.append(eps.clone())
            latents = sched.step(eps, t, latents)
            if trace is not None:
                trace.latents.append(latents.clone())
        return latents

    def _decode(self, latents, trace):
        img = self.vae.decode(latents / self.vae.cfg.scaling_factor)
        if trace is not None:
            trace.final_latents = latents.clone()
            trace.decoded = img.clone()
        return postprocess_image(img)

    @staticmethod
    def _embeds(prompt_embeds, negative_prompt_embeds, B, do_cfg):
        pe = prompt_embeds.repeat(B, 1, 1)
        if do_cfg:
            pe = torch.cat([negative_prompt_embeds.repeat(B, 1, 1), pe])
        return pe

    @torch.no_grad()
    def img2img(self, image: torch.Tensor, prompt_embeds, negative_prompt_embeds, *, strength=0.8,
                num_inference_steps=50, guidance_scale=7.5, generator=None, trace: Trace | None = None):
        """``image``: preprocessed f32 [B,3,H,W] in [-1,1].  Returns uint8 [B,H,W,3]."""
        if strength < 0 or strength > 1:
            raise ValueError(f"The value of strength should in [0.0, 1.0] but is {strength}")
        B = image.shape[0]
        do_cfg = guidance_scale > 1.0
        embeds = self._embeds(prompt_embeds, negative_prompt_embeds, B, do_cfg)
        sched = self._scheduler()
        sched.set_timesteps(num_inference_steps)
        timesteps, n = get_timesteps(sched, num_inference_steps, strength)
        if n < 1:
            raise ValueError("After adjusting the num_inference_steps by strength the number of steps is < 1")
        latent_t = timesteps[:1].repeat(B)
        init = self.vae.encode(image).sample(generator) * self.vae.cfg.scaling_factor       # RNG #1
        noise = randn_tensor(init.shape, generator, init.device, init.dtype)                 # RNG #2
        latents = sched.add_noise(init, noise, latent_t)
        if trace is not None:
            trace.init_latents = latents.clone()
        latents = self._loop(latents, timesteps, sched, embeds, guidance_scale, None, trace)
        return self._decode(latents, trace)

    @torch.no_grad()
    def inpaint(self, image: torch.Tensor, mask: torch.Tensor, prompt_embeds, negative_prompt_embeds, *,
                strength=1.0, num_inference_steps=50, guidance_scale=7.5, generator=None,
                trace: Trace | None = None):
        """``image`` f32 [B,3,512,512]; ``mask`` binarised f32 [B,1,512,512] (1 = fill)."""
        if strength < 0 or strength > 1:
            raise ValueError(f"The value of strength should in [0.0, 1.0] but is {strength}")
        B = image.shape[0]
        do_cfg = guidance_scale > 1.0
        embeds = self._embeds(prompt_embeds, negative_prompt_embeds, B, do_cfg)
        sched = self._scheduler()
        sched.set_timesteps(num_inference_steps)
        timesteps, n = get_timesteps(sched, num_inference_steps, strength)
        if n < 1:
            raise ValueError("After adjusting the num_inference_steps by strength the number of steps is < 1")
        latent_t = timesteps[:1].repeat(B)
        is_strength_max = strength == 1.0
        sf = self.vae.cfg.scaling_factor
        h, w = image.shape[2] // VAE_SCALE, image.shape[3] // VAE_SCALE
        shape = (B, 4, h, w)
        if not is_strength_max:
            image_latents = self.vae.encode(image).sample(generator) * sf                   # RNG #1
        noise = randn_tensor(shape, generator, image.device, image.dtype)                    # RNG #2
        latents = noise * sched.init_noise_sigma if is_strength_max else \
            sched.add_noise(image_latents, noise, latent_t)
        masked_image = image * (mask < 0.5)
        mask_lat = F.interpolate(mask, size=(h, w))                                          # nearest
        masked_lat = self.vae.encode(masked_image).sample(generator) * sf                    # RNG #3
        if do_cfg:
            mask_lat, masked_lat = torch.cat([mask_lat] * 2), torch.cat([masked_lat] * 2)
        if 4 + mask_lat.shape[1] + masked_lat.shape[1] != self.unet.cfg.in_channels:
            raise ValueError("Incorrect configuration settings: unet in_channels mismatch")
        if trace is not None:
            trace.init_latents = latents.clone()
        latents = self._loop(latents, timesteps, sched, embeds, guidance_scale, (mask_lat, masked_lat), trace)
        return self._decode(latents, trace)
