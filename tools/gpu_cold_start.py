"""Cold start of one pipeline (weights drawn on the device, repacked there, first call incl. graph capture) and a small
config-5 sweep on one GPU with the (psnr, ssim, lpips) triple.   python tools/gpu_cold_start.py [n_images]"""
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import sweep, synth
from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.cuda.init()
t0 = time.time()
pipe = StableDiffusionImg2ImgPipeline.from_random_init(seed=0, device="cuda")
torch.cuda.synchronize(); t1 = time.time()
pipe = pipe.to("cuda")
torch.cuda.synchronize(); t2 = time.time()
u8 = synth.batch("denoise", range(1))["input"]
kw = dict(prompt="clean high quality photo, no noise, sharp details", image=u8, strength=0.5, num_inference_steps=20,
          guidance_scale=5.0, output_type="np_u8")
pipe(generator=torch.Generator(device="cuda").manual_seed(42), **kw)
torch.cuda.synchronize(); t3 = time.time()
pipe(generator=torch.Generator(device="cuda").manual_seed(42), **kw)
torch.cuda.synchronize(); t4 = time.time()
print(f"cold start: random init {t1 - t0:.2f} s, .to(cuda) {t2 - t1:.2f} s, first call {t3 - t2:.2f} s, second call {t4 - t3:.3f} s", flush=True)
del pipe
torch.cuda.empty_cache()
t0 = time.time()
res = sweep.run_sweep(n_images=n)
print(f"sweep of {n} images x 4 tasks on one GPU: {time.time() - t0:.1f} s wall")
for task, r in res.items():
    print(task, r["seconds_rank0"], {k: round(float(v["mean"]), 4) for k, v in r["metrics"].items()})
