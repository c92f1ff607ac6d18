"""Sharded predict + evaluate sweep (BASELINE.json config 5): images x tasks across the GPUs of one box.

The reference runs this as two sequential scripts -- ``scripts/generate_predictions.py:66-88`` (a plain for-loop,
one ``pipeline.process`` per image) and ``scripts/evaluate_model.py:63-106`` (per-image PSNR/SSIM/LPIPS then
statistics).  Every image is an independent sampling run, so the work shards by image: rank r takes the work items
``i % world == r`` of the index-sorted list, runs them in batches through ``RestorationPipeline.process_batch``,
scores them on the host, and one all-gather of per-image float64 values hands rank 0 the complete, index-ordered
lists, which it aggregates with the reference's statistics.  No collective sits inside the sampling loop.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch
from PIL import Image

from . import metrics, synth
from .inference import RestorationPipeline

TASKS = ("denoise", "sr", "colorize", "inpaint")
BATCH = {"denoise": 8, "sr": 8, "colorize": 8, "inpaint": 8}


def shard(n_items: int, rank: int, world: int) -> list[int]:
    """Round-robin by global index (``i % world == rank``)."""
    return list(range(rank, n_items, world))


def run_task(pipe: RestorationPipeline, task: str, n_images: int, rank: int = 0, world: int = 1,
             size: int = 512, batch: int | None = None) -> tuple[list[int], dict[str, list[float]], float]:
    calc = metrics.MetricsCalculator(use_lpips=False)
    mine = shard(n_images, rank, world)
    bsz = batch or BATCH[task]
    vals: dict[str, list[float]] = {"psnr": [], "ssim": []}
    t0 = time.time()
    for s in range(0, len(mine), bsz):
        idx = mine[s:s + bsz]
        data = synth.batch(task, idx, size, size)
        ims = [Image.fromarray(a) for a in data["input"]]
        masks = [Image.fromarray(m) for m in data["mask"]] if "mask" in data else None
        outs = pipe.process_batch(ims, task, masks=masks)
        for o, gt in zip(outs, data["gt"]):
            m = calc.calculate_all(np.array(o.convert("RGB")), gt)
            vals["psnr"].append(m["psnr"])
            vals["ssim"].append(m["ssim"])
    return mine, vals, time.time() - t0


def run_sweep(n_images: int = 100, tasks=TASKS, size: int = 512, seed: int = 42, random_init: int = 0) -> dict | None:
    """Call from every rank (torchrun).  Returns the evaluation dict on rank 0, None elsewhere."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    cfg = {t: {"fine_tuned_dir": "nonexistent", "pretrained_id": "", "random_init": random_init + (1000 if t == "inpaint" else 0)}
           for t in TASKS}
    pipe = RestorationPipeline(device="cuda", config=cfg, seed=seed, strict=True)
    results = {}
    for task in tasks:
        idx, vals, secs = run_task(pipe, task, n_images, rank, world, size)
        full = metrics.gather_per_image(idx, vals)
        if rank == 0:
            results[task] = metrics.summarize(task, full, n_images)
            results[task]["seconds_rank0"] = secs
    return results if rank == 0 else None


if __name__ == "__main__":
    import json
    import torch.distributed as dist
    if "RANK" in os.environ:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    res = run_sweep(n_images=int(os.environ.get("SWEEP_IMAGES", "16")))
    if res is not None:
        print(json.dumps(res, default=float))
    if dist.is_initialized():
        dist.destroy_process_group()
