#!/usr/bin/env python3
"""Benchmark of the RestoraGen sampling loop on B200 (contract: see the task brief / DESIGN.md section "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): SD-1.5 img2img colorization,
512x512 grayscale input, batch 8 per GPU, 30 PNDM steps at strength 0.75 (= 23 UNet evaluations of batch 16 under
classifier-free guidance 7.5), one VAE encode and one VAE decode per image, random-init weights, synthetic inputs.
One "step" = one complete sampling run of that batch.  metric = 512x512 images per second, whole job.

  value   device-resident: uint8 inputs already in HBM, uint8 outputs left in HBM, CUDA-event timed
  e2e     through the public API (StableDiffusionImg2ImgPipeline.__call__) with HOST uint8 buffers in and out,
          host<->device copies and the final synchronisation inside the timed region
  roofline   the dominant kernel (conv_gemm_kernel, tcgen05 implicit GEMM): algorithmic FLOPs of every launch
             of one UNet evaluation / CUDA-event time of those launches, against MEASURED_PEAKS.json
  cpu_baseline   the fp32 oracle restatement of the reference path on the host cores (bounded sample)

--impl reference times the reference's own CPU path (the oracle port: diffusers is not installable offline) and
prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "512x512 colorize images/sec (SD-1.5 img2img, 30 PNDM steps, strength 0.75, CFG 7.5)"
UNIT = "images/s"
BATCH = 8
STEPS, STRENGTH, GUIDANCE = 30, 0.75, 7.5
UNET_FWD_PER_IMG = 46                       # 23 timesteps x 2 (CFG)
TFLOP_PER_IMG = 40.58
TFLOP_UNET_SAMPLE_FWD = 0.8033
WORKLOAD = {"workload": "SD-1.5 img2img colorization 512x512, batch 8/GPU, 30 PNDM steps (23 run), strength 0.75, "
                        "CFG 7.5, random-init UNet/VAE, synthetic gray inputs",
            "batch_per_gpu": BATCH, "unet_evals_per_step": 23, "unet_batch": 2 * BATCH,
            "l2": "no explicit flush: one step streams >10 GB of activations and 1.9 GB of weights through the "
                  "126 MB L2"}


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"burst": d["bf16_tflops"], "sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_unet_pair_seconds(reps: int = 2) -> tuple[float, int]:
    """Seconds for ONE classifier-free-guidance UNet evaluation of one image (2 sample-forwards, 64x64 latents) with
    the fp32 oracle on all host cores."""
    import torch
    from oracle.unet import UNet2DConditionModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.no_grad():
        m = UNet2DConditionModel().eval()
        g = torch.Generator().manual_seed(0)
        x = torch.randn((2, 4, 64, 64), generator=g)
        ctx = torch.randn((2, 77, 768), generator=g)
        m(x, torch.tensor(500.0), ctx)                      # warm-up (oneDNN primitive creation)
        t0 = time.time()
        for _ in range(reps):
            m(x, torch.tensor(500.0), ctx)
        return (time.time() - t0) / reps, cores


def cpu_images_per_sec(pair_s: float) -> float:
    # the UNet pair is 2 x 0.8033 TFLOP of the image's 40.58 TFLOP; scale by algorithmic work
    return 1.0 / (pair_s * TFLOP_PER_IMG / (2 * TFLOP_UNET_SAMPLE_FWD))


CPU_SAMPLE = ("one CFG UNet evaluation of one image (2 sample-forwards at 64x64 latents, 1.607 TFLOP, fp32 oracle, "
              "torch CPU) timed and scaled by algorithmic work to the image's 40.58 TFLOP (46 sample-forwards + "
              "VAE encode + decode)")


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times = []
    pair_s, cores = None, os.cpu_count() or 1
    for i in range(max(1, min(args.steps, 3))):
        pair_s, cores = cpu_unet_pair_seconds(reps=1)
        times.append(pair_s)
    pair_s = sum(times) / len(times)
    v = cpu_images_per_sec(pair_s)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": WORKLOAD,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": CPU_SAMPLE},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = fp32 oracle restatement of the diffusers path on host cores (diffusers itself is not "
                    "installable offline; see DESIGN.md)"}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist
    from image_restoration_and_enhancement_b200 import ops, synth
    from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # the contract is ONE JSON line on stdout: libraries that chat on stdout (NCCL prints its version banner there when
    # NCCL_DEBUG=VERSION is set on the box) are sent to stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pipe = StableDiffusionImg2ImgPipeline.from_random_init(seed=0).to(dev)
    prompt = "vibrant realistic natural colors, colorful, high quality photo, detailed, full color, rich colors"
    # every rank works on its own slice of the (synthetic) image stream: weak scaling, no data-path collective
    host_u8 = torch.from_numpy(synth.batch("colorize", range(rank * BATCH, (rank + 1) * BATCH))["input"]).pin_memory()
    dev_u8 = host_u8.to(dev)

    def gens():
        return [torch.Generator(device=dev).manual_seed(42) for _ in range(BATCH)]

    def step_device():
        return pipe(prompt=prompt, image=dev_u8, strength=STRENGTH, num_inference_steps=STEPS,
                    guidance_scale=GUIDANCE, generator=gens(), output_type="u8_device").images

    def step_e2e():
        return pipe(prompt=prompt, image=host_u8.numpy(), strength=STRENGTH, num_inference_steps=STEPS,
                    guidance_scale=GUIDANCE, generator=gens(), output_type="np_u8").images

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = pipe.launches()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k, pipe.launches() - l0

    for _ in range(max(args.warmup, 3) if args.warmup >= 0 else 3):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, _ = timed(step_e2e, max(1, min(args.steps, 3)))

    roof = cpu = None
    if rank == 0:
        roof = conv_roofline(pipe, dev)
        if world == 1 and not args.no_cpu:
            pair_s, cores = cpu_unet_pair_seconds(reps=1)
            cpu = {"value": cpu_images_per_sec(pair_s), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": CPU_SAMPLE, "unet_cfg_pair_seconds": pair_s}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank != 0:
        return 0
    total = BATCH * world
    line = {"metric": METRIC, "value": total / (ms_dev / 1000.0), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": {**WORKLOAD, "parallelism": f"dp{world}"},
            "e2e": {"value": total / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(host_u8.numel()),
                    "d2h_bytes_per_step": int(host_u8.numel()), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "achieved_tflops_whole_step": TFLOP_PER_IMG * BATCH / (ms_dev / 1000.0),
            "frac_of_bf16_sustained_whole_step": TFLOP_PER_IMG * BATCH / (ms_dev / 1000.0) / peaks()["sustained"]}
    print(json.dumps(line))
    return 0


def conv_roofline(pipe, dev) -> dict:
    """Per-launch CUDA-event timing of every conv_gemm_kernel launch of one eager UNet evaluation (batch 16)."""
    import torch
    from image_restoration_and_enhancement_b200 import ops
    pk = peaks()
    unet = pipe._unet
    lat = torch.randn((BATCH, 64, 64, 4), device=dev)
    ts = torch.full((2 * BATCH,), 500.0, device=dev)
    unet.forward(lat, ts)
    torch.cuda.synchronize()
    ops.PROFILE = []
    unet.forward(lat, ts)
    torch.cuda.synchronize()
    allrecs, ops.PROFILE = ops.PROFILE, None
    recs = [r for r in allrecs if r[3] == "gemm"]
    att = [r for r in allrecs if r[3] == "attention"]
    tot_ms = sum(r[0].elapsed_time(r[1]) for r in recs)
    tot_flop = sum(r[2] for r in recs)
    att_ms = sum(r[0].elapsed_time(r[1]) for r in att)
    att_flop = sum(r[2] for r in att)
    achieved = tot_flop / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
    traffic = None
    tp = ROOT / "profiles" / "conv_gemm_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
    return {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, all launches of one UNet evaluation)",
            "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
            "peak_kind": f"bf16 sustained, {pk['source']}", "frac_of_burst": achieved / pk["burst"],
            "launches_timed": len(recs), "ms_total": tot_ms, "traffic": traffic,
            "attention_kernel": {"achieved": att_flop / (att_ms / 1e3) / 1e12 if att_ms > 0 else 0.0, "unit": "TFLOP/s",
                                 "launches_timed": len(att), "ms_total": att_ms}}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
