"""Per-launch profile of one eager UNet evaluation (UNet batch 2*B) and of VAE encode / decode (batch B) on the GPU box,
plus the phase split of one graph-replayed sampling run.  CUDA-event timed per op (not under a profiler).

    python tools/gpu_layer_profile.py [tag] [B]      -> gpurun_out/layers_<tag>.txt / .json   (B=8: colorize, B=1: denoise)
"""
import collections
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import os

import torch

from image_restoration_and_enhancement_b200 import _lib
if os.environ.get("RG_LIB"):            # another build of the library, for same-box A/B runs
    _lib.LIB_PATH = Path(os.environ["RG_LIB"]).resolve()
from image_restoration_and_enhancement_b200 import ops, synth
from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if "nosplit" in sys.argv[3:]:
    ops.SPLITK = False
if "nopdl" in sys.argv[3:]:
    _lib.load().rg_set_pdl(0)
dev = torch.device("cuda", 0)
pipe = StableDiffusionImg2ImgPipeline.from_random_init(seed=0, device="cuda").to(dev)
out_lines = []


def say(s=""):
    print(s, flush=True)
    out_lines.append(s)


def profile(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        ops.PROFILE = []
        fn()
        torch.cuda.synchronize()
        recs, ops.PROFILE = ops.PROFILE, None
        rows = [(r[0].elapsed_time(r[1]) * 1e3, r[2], r[3], r[4]) for r in recs]
        if best is None or sum(r[0] for r in rows) < sum(r[0] for r in best):
            best = rows
    agg = collections.OrderedDict()
    for us, work, kind, desc in best:
        a = agg.setdefault((kind, desc), [0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += work
    tot = sum(r[0] for r in best)
    say(f"==== {name}: {len(best)} profiled ops, {tot / 1e3:.3f} ms (sum of per-op event times, eager)")
    bykind = collections.defaultdict(lambda: [0.0, 0.0])
    for (kind, desc), (n, us, work) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rate = work / us / 1e6 if kind in ("gemm", "attention") else work / us / 1e3      # TFLOP/s | GB/s
        unit = "TF/s" if kind in ("gemm", "attention") else "GB/s"
        say(f"{us:9.1f} us {100 * us / tot:5.1f}%  x{n:<3d} {us / n:8.1f} us/launch {rate:8.1f} {unit}  {kind:10s} {desc}")
        bykind[kind][0] += us; bykind[kind][1] += work
    for kind, (us, work) in bykind.items():
        rate = work / us / 1e6 if kind in ("gemm", "attention") else work / us / 1e3
        say(f"  -- {kind:10s} {us / 1e3:8.3f} ms  {100 * us / tot:5.1f}%   {rate:8.1f} {'TF/s' if kind in ('gemm', 'attention') else 'GB/s'}")
    return {"name": name, "total_ms": tot / 1e3,
            "rows": [{"kind": k, "desc": d, "n": n, "us": us, "work": w} for (k, d), (n, us, w) in agg.items()]}


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
with torch.no_grad():
    unet, vae = pipe._unet, pipe._vae
    ctx = torch.randn((2 * B, 77, 768), device=dev)
    unet.prepare_context(ctx)
    lat = torch.randn((B, 64, 64, 4), device=dev)
    ts = torch.full((2 * B,), 500.0, device=dev)
    res["unet"] = profile(f"UNet evaluation, batch {2 * B} ({B} images x CFG)", lambda: unet.forward(lat, ts))
    img = torch.rand((B, 512, 512, 3), device=dev) * 2 - 1
    res["vae_enc"] = profile(f"VAE encode, {B} images", lambda: vae.encode_moments(img))
    res["vae_dec"] = profile(f"VAE decode, {B} images", lambda: vae.decode(lat * 0.18215))

    # whole-op timings (eager back-to-back and graph replay)
    say()
    say(f"eager UNet eval b{2 * B}: {timed(lambda: unet.forward(lat, ts)):.3f} ms")
    run = pipe._unet_step_fn(lat, 2 * B); trow = unet.time_embedding(torch.full((1,), 500.0, device=dev))
    say(f"graph UNet eval b{2 * B}: {timed(lambda: run(trow)):.3f} ms")
    say(f"VAE encode b{B}: {timed(lambda: vae.encode_moments(img)):.3f} ms")
    say(f"VAE decode b{B}: {timed(lambda: vae.decode(lat * 0.18215)):.3f} ms")
    task = "colorize" if B > 1 else "denoise"
    u8 = torch.from_numpy(synth.batch(task, range(B))["input"]).to(dev)
    gens = lambda: [torch.Generator(device=dev).manual_seed(42) for _ in range(B)]
    if B > 1:
        full = lambda: pipe(prompt="vibrant realistic natural colors, colorful, high quality photo, detailed, full color, rich colors",
                            image=u8, strength=0.75, num_inference_steps=30, guidance_scale=7.5, generator=gens(),
                            output_type="u8_device")
    else:
        full = lambda: pipe(prompt="clean high quality photo, no noise, sharp details", image=u8, strength=0.5,
                            num_inference_steps=20, guidance_scale=5.0, generator=gens(), output_type="u8_device")
    say(f"full {task} run b{B}: {timed(full, 3):.3f} ms")
    # host-side view of the same run: wall clock per call incl. Python / ctypes / tensor-map encodes of the eager VAE
    import time
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        full()
    t_issue = (time.time() - t0) / 3
    torch.cuda.synchronize()
    say(f"full {task} run b{B}: host issue time {t_issue * 1e3:.3f} ms per call (before the final sync)")

(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"layers_{tag}.txt").write_text("\n".join(out_lines) + "\n")
(ROOT / "gpurun_out" / f"layers_{tag}.json").write_text(json.dumps(res))
