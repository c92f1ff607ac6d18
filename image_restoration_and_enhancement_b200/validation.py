"""Training-time validation caller (SURVEY 8f "f4", the reference's fifth caller of the sampling loop).

``scripts/train_denoising.py:328-520`` ``run_validation`` takes the diffusers pipeline object, swaps in the UNet being
trained (``pipeline.unet = accelerator.unwrap_model(unet_model)``), samples ``num_samples`` evenly spaced validation
items with a fixed prompt (strength 0.3, 20 steps, guidance 5.0), scores them (RGB PSNR / SSIM, Y-channel PSNR / SSIM,
per-sigma buckets), saves input | result | ground-truth strips and returns the means.  This is that function against
the B200 pipeline classes: the assignment reloads the kernels' weights from the module's ``state_dict()``
(``pipelines._SDPipelineBase.unet`` setter) and the metrics run on the GPU with the same float64 bits as the CPU path.
Accelerate is not needed: ``unet_model`` is used as given.
"""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import torch
from PIL import Image

from .metrics import MetricsCalculator

logger = logging.getLogger(__name__)
VALIDATION_PROMPT = "a photograph, high quality, detailed, sharp"          # scripts/train_denoising.py:399


def _to_u8(t: torch.Tensor) -> np.ndarray:
    """[-1, 1] CHW tensor -> uint8 HWC, as ``:395-397`` / ``:411-413`` (clamp, * 255, truncate)."""
    vis = torch.clamp((t.detach().float().cpu() + 1.0) / 2.0, 0, 1)
    return (vis.permute(1, 2, 0).numpy() * 255).astype(np.uint8)


def run_validation(epoch: int, val_dataset, pipeline, unet_model, output_dir, num_samples: int = 4,
                   device: str = "cuda", prompt: str = VALIDATION_PROMPT, strength: float = 0.3,
                   num_inference_steps: int = 20, guidance_scale: float = 5.0, seed: int | None = None) -> dict | None:
    """Returns {'psnr', 'ssim', 'psnr_y', 'ssim_y', 'by_sigma', 'num_samples'} or None if there is nothing to validate.
    ``seed``: the reference draws its noise from the global RNG (no generator is passed, ``:400-406``), so its validation
    numbers are not reproducible; with a seed every sample gets ``torch.Generator(device).manual_seed(seed)``."""
    import cv2
    if val_dataset is None or len(val_dataset) == 0:
        logger.warning("Validation dataset is None or empty, skipping validation")
        return None
    val_dir = Path(output_dir) / "val_samples"
    val_dir.mkdir(parents=True, exist_ok=True)
    num_samples = min(num_samples, len(val_dataset))
    sample_indices = np.linspace(0, len(val_dataset) - 1, num_samples, dtype=int)
    calc = MetricsCalculator(use_lpips=False, use_fid=False, device=device)
    if unet_model is not None:
        pipeline.unet = unet_model                                          # :351
    pipeline = pipeline.to(device)                                          # :353
    for attr, val in (("safety_checker", None), ("feature_extractor", None), ("requires_safety_checker", False)):
        if hasattr(pipeline, attr):
            setattr(pipeline, attr, val)                                    # :355-360
    pipeline.unet.eval(); pipeline.vae.eval(); pipeline.text_encoder.eval()  # :362-364
    psnrs, ssims, psnrs_y, ssims_y = [], [], [], []
    buckets: dict[int, dict[str, list]] = {}
    try:
        with torch.no_grad():
            for i, idx in enumerate(sample_indices):
                sample = val_dataset[int(idx)]
                input_np, gt_np = _to_u8(sample["input"]), _to_u8(sample["gt"])
                extra = {} if seed is None else {"generator": torch.Generator(device=device).manual_seed(seed)}
                result = pipeline(prompt=prompt, image=Image.fromarray(input_np), strength=strength,
                                  num_inference_steps=num_inference_steps, guidance_scale=guidance_scale, **extra).images[0]
                result_np = np.array(result)
                if result_np.sum() < 1000:
                    logger.warning(f"Sample {idx} produced dark output (sum={result_np.sum()})")
                if result_np.shape[:2] != gt_np.shape[:2]:
                    result_np = cv2.resize(result_np, (gt_np.shape[1], gt_np.shape[0]))
                m = calc.calculate_all(result_np, gt_np)
                psnrs.append(m["psnr"]); ssims.append(m["ssim"])
                # Y channel of YCrCb (:368-383)
                my = calc.calculate_all(cv2.cvtColor(result_np, cv2.COLOR_RGB2YCrCb)[:, :, 0],
                                        cv2.cvtColor(gt_np, cv2.COLOR_RGB2YCrCb)[:, :, 0])
                psnrs_y.append(my["psnr"]); ssims_y.append(my["ssim"])
                sigma = sample.get("sigma") if hasattr(sample, "get") else None
                if sigma is not None:
                    b = buckets.setdefault(int(round(float(sigma))), {"psnr": [], "ssim": [], "psnr_y": [], "ssim_y": []})
                    b["psnr"].append(m["psnr"]); b["ssim"].append(m["ssim"])
                    b["psnr_y"].append(my["psnr"]); b["ssim_y"].append(my["ssim"])
                if input_np.shape[:2] != gt_np.shape[:2]:
                    input_np = cv2.resize(input_np, (gt_np.shape[1], gt_np.shape[0]))
                Image.fromarray(np.hstack([input_np, result_np, gt_np])).save(
                    val_dir / f"epoch_{epoch + 1}_sample_{i + 1}_idx{idx}.png")
    finally:
        pipeline.unet.train()                                               # :466
    if not psnrs:
        return None
    return {"psnr": float(np.mean(psnrs)), "ssim": float(np.mean(ssims)), "psnr_y": float(np.mean(psnrs_y)),
            "ssim_y": float(np.mean(ssims_y)), "num_samples": len(psnrs),
            "by_sigma": {k: {m: float(np.mean(v)) for m, v in b.items() if v} for k, b in sorted(buckets.items())},
            "per_image": {"psnr": psnrs, "ssim": ssims}}
