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


# sweep task name -> directory name of the reference's pairs/ and predictions/ trees
TASK_DIR = {"denoise": "denoise", "sr": "sr_x4", "colorize": "colorize", "inpaint": "inpaint"}


def _score(preds: list[np.ndarray], gts: list[np.ndarray], backend: str, lpips_model=None) -> dict[str, list[float]]:
    """Per-image PSNR / SSIM (/ LPIPS).  "gpu": csrc/metrics.cu (+ lpips.LPIPSB200) on the uploaded u8 batch; "cpu": the
    numpy/scipy bookkeeping.  PSNR / SSIM give the same float64 bits either way (tests/test_kernels_gpu.py::metrics_*);
    LPIPS exists on the GPU only."""
    preds = [metrics._match_shape(p, g) for p, g in zip(preds, gts)]
    if backend == "gpu":
        out: dict[str, list] = {"psnr": [None] * len(preds), "ssim": [None] * len(preds)}
        if lpips_model is not None:
            out["lpips"] = [None] * len(preds)
        by_shape: dict[tuple, list[int]] = {}
        for i, g in enumerate(gts):
            by_shape.setdefault(g.shape, []).append(i)
        for idx in by_shape.values():
            pd = torch.from_numpy(np.stack([preds[i] for i in idx])).cuda()
            gd = torch.from_numpy(np.stack([gts[i] for i in idx])).cuda()
            p, s = metrics.psnr_ssim_device(pd, gd)
            lp = lpips_model(pd, gd) if lpips_model is not None else None
            for j, i in enumerate(idx):
                out["psnr"][i], out["ssim"][i] = p[j], s[j]
                if lp is not None:
                    out["lpips"][i] = lp[j]
        return out
    if backend != "cpu":
        raise ValueError(f"unknown metrics backend {backend}")
    calc = metrics.MetricsCalculator(use_lpips=False)
    return {"psnr": [calc.calculate_psnr(p, g) for p, g in zip(preds, gts)],
            "ssim": [calc.calculate_ssim(p, g) for p, g in zip(preds, gts)]}


def run_task(pipe, task: str, n_images: int, rank: int = 0, world: int = 1, size: int = 512, batch: int | None = None,
             handoff_mode: str = "memory", metrics_backend: str = "gpu", workdir=None, overlap: bool = True,
             lpips_model=None) -> tuple[list[int], dict[str, list[float]], float]:
    """Predict + score this rank's share of one task.

    ``handoff_mode`` selects what sits between the pipeline and the metrics (see handoff.py):
      "disk"   the reference's two-script flow: the synthetic pairs are written under ``workdir/pairs`` with
               cv2.imwrite, inputs are re-read with PIL, predictions saved by name under ``workdir/predictions`` and
               re-read, like the ground truth, with cv2.imread;
      "memory" the same encoders / decoders on byte buffers (identical arrays, no file system);
      "none"   raw arrays straight through (no codec; NOT what the reference's evaluator scores for .jpg tasks).
    ``overlap``: prepare batch k+1 and score batch k-1 on host threads while the GPU samples batch k.
    """
    from concurrent.futures import ThreadPoolExecutor
    from pathlib import Path
    from . import handoff
    mine = shard(n_images, rank, world)
    bsz = batch or BATCH[task]
    tdir = TASK_DIR[task]
    if handoff_mode not in ("none", "memory", "disk"):
        raise ValueError(f"unknown handoff_mode {handoff_mode}")
    if handoff_mode == "disk":
        if workdir is None:
            raise ValueError('handoff_mode="disk" needs workdir')
        pair_root, pred_dir = Path(workdir) / "pairs", Path(workdir) / "predictions" / tdir / "test"
        pred_dir.mkdir(parents=True, exist_ok=True)
    device = torch.cuda.current_device() if (metrics_backend == "gpu" and torch.cuda.is_available()) else None

    def prepare(idx):
        """Host side of one batch before the sampling run: synthesis + the dataset codecs (no CUDA)."""
        items = [synth.make_pair(task, i, size, size) for i in idx]
        names = [handoff.input_name(tdir, i) for i in idx]
        masks = None
        if handoff_mode == "none":
            ims = [Image.fromarray(it["input"]) for it in items]
            gts = [it["gt"] for it in items]
            if "mask" in items[0]:
                masks = [Image.fromarray(it["mask"]) for it in items]
        elif handoff_mode == "memory":
            gray = tdir == "colorize"
            ims = [handoff.roundtrip_dataset_image(it["input"][:, :, 0] if gray else it["input"], nm, "pil")
                   for it, nm in zip(items, names)]
            gts = [handoff.roundtrip_dataset_image(it["gt"], handoff.gt_name(tdir, i), "cv2") for it, i in zip(items, idx)]
            if "mask" in items[0]:
                masks = [handoff.roundtrip_dataset_image(it["mask"], nm, "pil_l") for it, nm in zip(items, names)]
        else:
            for it, i in zip(items, idx):
                handoff.write_pairs(pair_root, tdir, i, it)
            base = pair_root / tdir / "test"
            ims = [Image.open(base / "input" / nm).convert("RGB") for nm in names]       # generate_predictions.py:68
            gts = [handoff.load_image(base / "gt" / handoff.gt_name(tdir, i)) for i in idx]
            if "mask" in items[0]:
                masks = [Image.open(base / "mask" / nm).convert("L") for nm in names]    # generate_predictions.py:74
        return names, ims, gts, masks

    def score(outs, names, gts):
        """Host side after the sampling run: the prediction hand-off, then the metrics."""
        if device is not None:
            torch.cuda.set_device(device)
        if handoff_mode == "none":
            preds = [np.array(o.convert("RGB")) for o in outs]
        elif handoff_mode == "memory":
            preds = [handoff.roundtrip_prediction(o, nm) for o, nm in zip(outs, names)]
        else:
            for o, nm in zip(outs, names):
                handoff.save_prediction(o, pred_dir / nm)                                # generate_predictions.py:83-84
            preds = [handoff.load_image(pred_dir / nm) for nm in names]
        return _score(preds, gts, metrics_backend, lpips_model)

    batches = [mine[s:s + bsz] for s in range(0, len(mine), bsz)]
    vals: dict[str, list[float]] = {"psnr": [], "ssim": []}
    if lpips_model is not None and metrics_backend == "gpu":
        vals["lpips"] = []

    def collect(d):
        for k, v in d.items():
            vals[k] += v
    t0 = time.time()
    if not overlap or len(batches) < 2:
        for idx in batches:
            names, ims, gts, masks = prepare(idx)
            collect(score(pipe.process_batch(ims, task, masks=masks), names, gts))
        return mine, vals, time.time() - t0
    # Software pipeline over batches: while the GPU samples batch k, one host thread prepares batch k+1 (synthesis,
    # codecs) and another scores batch k-1 (codec round trip, metric kernels on the default stream).  PIL / OpenCV /
    # numpy release the GIL, so the three stages really overlap; results are collected in batch order.
    with ThreadPoolExecutor(max_workers=2) as pool:
        pending = []
        nxt = pool.submit(prepare, batches[0])
        last_n = None
        for k, idx in enumerate(batches):
            names, ims, gts, masks = nxt.result()
            if k + 1 < len(batches):
                nxt = pool.submit(prepare, batches[k + 1])
            if last_n is not None and len(idx) != last_n:
                for f in pending:            # a new batch size captures a new CUDA graph: no concurrent CUDA work then
                    f.result()
            last_n = len(idx)
            outs = pipe.process_batch(ims, task, masks=masks)
            pending.append(pool.submit(score, outs, names, gts))
        for f in pending:
            collect(f.result())
    return mine, vals, time.time() - t0


def run_sweep(n_images: int = 100, tasks=TASKS, size: int = 512, seed: int = 42, random_init: int = 0,
              handoff_mode: str = "memory", metrics_backend: str = "gpu", workdir=None, use_lpips: bool = True,
              lpips_seed: int = 0) -> dict | None:
    """Call from every rank (torchrun).  Returns the evaluation dict on rank 0, None elsewhere.  ``use_lpips``: per-image
    LPIPS next to PSNR / SSIM (GPU metric backend; seeded random-init AlexNet, the same on every rank -- see lpips.py)."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    cfg = {t: {"fine_tuned_dir": "nonexistent", "pretrained_id": "", "random_init": random_init + (1000 if t == "inpaint" else 0)}
           for t in TASKS}
    pipe = RestorationPipeline(device="cuda", config=cfg, seed=seed, strict=True)
    lp = None
    if use_lpips and metrics_backend == "gpu":
        from .lpips import LPIPSB200, random_lpips_state_dict
        lp = LPIPSB200(random_lpips_state_dict(lpips_seed), device=f"cuda:{torch.cuda.current_device()}")
    results = {}
    from . import ops
    # The sweep's contract is that the statistics do not depend on how the images were sharded (1 process or 8, ragged last
    # batches): every kernel is batch-invariant bit for bit except the K split of the few-tile GEMMs, which follows the
    # batch size -- so the sweep samples without it (it only matters below batch 4; cost at batch 8: ~1 %).
    with ops.splitk(False):
        for task in tasks:
            idx, vals, secs = run_task(pipe, task, n_images, rank, world, size, handoff_mode=handoff_mode,
                                       metrics_backend=metrics_backend, workdir=workdir, lpips_model=lp)
            full = metrics.gather_per_image(idx, vals)
            if rank == 0:
                results[task] = metrics.summarize(task, full, n_images)
                results[task]["seconds_rank0"] = secs
                results[task]["images_per_s_wall"] = n_images / secs if secs > 0 else None
    return results if rank == 0 else None


if __name__ == "__main__":
    import json
    import torch.distributed as dist
    if "RANK" in os.environ:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    t_all = time.time()
    res = run_sweep(n_images=int(os.environ.get("SWEEP_IMAGES", "16")), handoff_mode=os.environ.get("SWEEP_HANDOFF", "memory"),
                    metrics_backend=os.environ.get("SWEEP_METRICS", "gpu"), workdir=os.environ.get("SWEEP_WORKDIR"))
    if res is not None:
        res["_wall_seconds_incl_model_setup"] = time.time() - t_all
        res["_world_size"] = dist.get_world_size() if dist.is_initialized() else 1
        print(json.dumps(res, default=float))
    if dist.is_initialized():
        dist.destroy_process_group()
