"""Where the first (cold) pipeline call of a process goes: text encoder, VAE / UNet graph captures, the rest.
    python tools/gpu_cold_call_profile.py"""
import gc
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from image_restoration_and_enhancement_b200 import synth
from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline

t0 = time.time()
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
print(f"CUDA context: {time.time() - t0:.2f} s")
from image_restoration_and_enhancement_b200 import pipelines as P, weights as W, unet as U, vae as V, text_encoder as T


def timed(mod, name):
    fn = getattr(mod, name)

    def w(*a, **k):
        torch.cuda.synchronize(); t = time.time()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        print(f"   {name}: {time.time() - t:.2f} s")
        return r
    setattr(mod, name, w)


for mod, name in ((P, "random_state_dict"), (P, "_RandomCLIPText"), (P, "UNetB200"), (P, "VAEB200"), (T, "CLIPTextB200")):
    timed(mod, name)
t0 = time.time()
pipe = StableDiffusionImg2ImgPipeline.from_random_init(seed=0, device="cuda")
torch.cuda.synchronize()
print(f"from_random_init: {time.time() - t0:.2f} s")
t0 = time.time()
pipe = pipe.to("cuda")
torch.cuda.synchronize()
print(f"to(cuda): {time.time() - t0:.2f} s")
acc = {}


def wrap(obj, name):
    fn = getattr(obj, name)

    def w(*a, **k):
        torch.cuda.synchronize(); t = time.time()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc.setdefault(name, []).append(time.time() - t)
        return r
    setattr(obj, name, w)


wrap(pipe, "_unet_step_fn"); wrap(pipe, "_vae_graphed")
if hasattr(pipe, "encode_prompt"): wrap(pipe, "encode_prompt")
if hasattr(pipe, "_encode_prompt"): wrap(pipe, "_encode_prompt")
t = time.time(); gc.collect(); print(f"gc.collect(): {time.time() - t:.3f} s")
for B in (8, 8, 5, 5):
    data = synth.batch("denoise", range(B))
    acc.clear()
    torch.cuda.synchronize(); t = time.time()
    pipe(prompt="clean high quality photo, no noise, sharp details", image=data["input"], strength=0.5, num_inference_steps=20,
         guidance_scale=5.0, generator=[torch.Generator(device="cuda").manual_seed(42) for _ in range(B)], output_type="np_u8")
    torch.cuda.synchronize()
    print(f"call B={B}: {time.time() - t:.3f} s; " + "; ".join(f"{k}: " + "+".join(f"{x:.3f}" for x in v) for k, v in acc.items()))
