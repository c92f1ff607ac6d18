"""Where one bench step spends its GPU time: CUDA events around the pipeline phases.
    python tools/gpu_step_breakdown.py [colorize|denoise] [batch]      (defaults: colorize 8; config 1 = denoise 1)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import ops, synth
from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline

dev = torch.device("cuda", 0)
pipe = StableDiffusionImg2ImgPipeline.from_random_init(seed=0, device="cuda").to(dev)
task = sys.argv[1] if len(sys.argv) > 1 else "colorize"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
u8 = torch.from_numpy(synth.batch(task, range(B))["input"]).to(dev)
if task == "colorize":
    prompt, kw = "vibrant realistic natural colors, colorful, high quality photo, detailed, full color, rich colors", dict(strength=0.75, num_inference_steps=30, guidance_scale=7.5)
else:
    prompt, kw = "clean high quality photo, no noise, sharp details", dict(strength=0.5, num_inference_steps=20, guidance_scale=5.0)
gens = lambda: [torch.Generator(device=dev).manual_seed(42) for _ in range(B)]
call = lambda: pipe(prompt=prompt, image=u8, generator=gens(), output_type="u8_device", **kw)
for _ in range(2): call()
torch.cuda.synchronize()

marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))

# instrument by wrapping the pipeline's internals
orig_enc, orig_dec, orig_loop, orig_prep = pipe._vae.encode_moments, pipe._vae.decode, pipe._sample_loop, pipe._unet.prepare_context
def enc(x): mark("encode>"); r = orig_enc(x); mark("encode<"); return r
def dec(x): mark("decode>"); r = orig_dec(x); mark("decode<"); return r
def loop(*a, **k): mark("loop>"); r = orig_loop(*a, **k); mark("loop<"); return r
def prep(c): mark("prep>"); r = orig_prep(c); mark("prep<"); return r
pipe._vae.encode_moments, pipe._vae.decode, pipe._sample_loop, pipe._unet.prepare_context = enc, dec, loop, prep
orig_vg = pipe._vae_graphed
def vg(kind, x):                      # small batches: the VAE runs as a CUDA graph replay (the wrappers above are not called)
    mark(f"vae-{kind}>"); r = orig_vg(kind, x); mark(f"vae-{kind}<"); return r
if B <= 3:
    pipe._vae.encode_moments, pipe._vae.decode = orig_enc, orig_dec
    pipe._vae_graphed = vg
for rep in range(2):
    marks.clear()
    mark("start"); t0 = time.perf_counter()
    call()
    mark("end"); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"wall {1e3 * (t1 - t0):.1f} ms")
for (n0, e0, h0), (n1, e1, h1) in zip(marks[:-1], marks[1:]):
    print(f"{n0:10s} -> {n1:10s}  gpu {e0.elapsed_time(e1):8.2f} ms   host {1e3 * (h1 - h0):8.2f} ms")
print(f"total gpu {marks[0][1].elapsed_time(marks[-1][1]):.2f} ms")
