"""Seeded synthetic inputs for the benchmark and the config-5 sweep.

There is no dataset on disk (``data/pairs`` is not shipped with the reference and there is no network), so clean
images are smooth seeded random fields and the four task inputs are produced with the degradation recipe of the
reference's ``scripts/make_synthetic_pairs.py``: Gaussian noise sigma ~ U(5, 8) (``:29-35,170``); Gaussian blur
k in {3,5,7} then bicubic /4 for sr (``:67-81``); LAB-L grayscale for colorize (``:84-90``); free-form stroke masks,
masked pixels set to 0, for inpaint (``:104-114,191-192``).  Everything is driven by ``numpy.random.default_rng``
so the same index always yields the same pair on every rank.
"""
from __future__ import annotations

import numpy as np


def clean_image(index: int, H: int = 512, W: int = 512) -> np.ndarray:
    """Smooth seeded RGB uint8 image: low-pass random field + sinusoid + vertical gradient."""
    import cv2
    rng = np.random.default_rng(1234 + index)
    low = rng.standard_normal((H // 32 + 2, W // 32 + 2, 3)).astype(np.float32)
    up = cv2.resize(low, (W, H), interpolation=cv2.INTER_CUBIC)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = 127.0 + 50.0 * up + 30.0 * np.sin(xx / 37.0)[..., None] + 20.0 * (yy / H)[..., None]
    return np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))


def make_pair(task: str, index: int, H: int = 512, W: int = 512, sr_scale: int = 4) -> dict:
    """Returns {"input": uint8 HWC, "gt": uint8 HWC[, "mask": uint8 HW]} for one work item."""
    import cv2
    rng = np.random.default_rng(977 * index + 13)
    gt = clean_image(index, H, W)
    if task == "denoise":
        sigma = rng.uniform(5, 8)
        noisy = gt.astype(np.float32) + rng.standard_normal(gt.shape).astype(np.float32) * sigma
        return {"input": np.clip(noisy, 0, 255).astype(np.uint8), "gt": gt}
    if task in ("sr", "super_resolution"):
        k = int(rng.choice([3, 5, 7]))
        lr = cv2.resize(cv2.GaussianBlur(gt, (k, k), sigmaX=0), (W // sr_scale, H // sr_scale), interpolation=cv2.INTER_CUBIC)
        # training-time validation upsamples bicubically to the target size before the img2img call
        # (scripts/train_super_resolution.py:386); the sweep feeds that to the pipeline
        return {"input": cv2.resize(lr, (W, H), interpolation=cv2.INTER_CUBIC), "gt": gt, "lr": lr}
    if task == "colorize":
        gray = cv2.cvtColor(gt, cv2.COLOR_RGB2LAB)[:, :, 0]
        return {"input": np.ascontiguousarray(np.stack([gray] * 3, axis=2)), "gt": gt}
    if task == "inpaint":
        mask = np.zeros((H, W), dtype=np.uint8)
        for _ in range(int(rng.integers(3, 8))):
            pts = [(int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(int(rng.integers(4, 9)))]
            th = int(rng.integers(5, 21))
            for a, b in zip(pts[:-1], pts[1:]):
                cv2.line(mask, a, b, color=255, thickness=th)
        masked = gt.copy()
        masked[mask == 255] = 0
        return {"input": masked, "gt": gt, "mask": mask}
    raise ValueError(f"unknown task {task}")


def batch(task: str, indices, H: int = 512, W: int = 512) -> dict:
    items = [make_pair(task, i, H, W) for i in indices]
    out = {"input": np.stack([it["input"] for it in items]), "gt": np.stack([it["gt"] for it in items])}
    if "mask" in items[0]:
        out["mask"] = np.stack([it["mask"] for it in items])
    return out
