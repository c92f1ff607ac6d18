"""Where one bench step (colorize, batch 8) spends its GPU time: CUDA events around the pipeline phases."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import ops, synth
from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline

dev = torch.device("cuda", 0)
pipe = StableDiffusionImg2ImgPipeline.from_random_init(seed=0).to(dev)
B = 8
u8 = torch.from_numpy(synth.batch("colorize", range(B))["input"]).to(dev)
prompt = "vibrant realistic natural colors, colorful, high quality photo, detailed, full color, rich colors"
gens = lambda: [torch.Generator(device=dev).manual_seed(42) for _ in range(B)]
call = lambda: pipe(prompt=prompt, image=u8, strength=0.75, num_inference_steps=30, guidance_scale=7.5, generator=gens(), output_type="u8_device")
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
for rep in range(2):
    marks.clear()
    mark("start"); t0 = time.perf_counter()
    call()
    mark("end"); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"wall {1e3 * (t1 - t0):.1f} ms")
for (n0, e0, h0), (n1, e1, h1) in zip(marks[:-1], marks[1:]):
    print(f"{n0:10s} -> {n1:10s}  gpu {e0.elapsed_time(e1):8.2f} ms   host {1e3 * (h1 - h0):8.2f} ms")
print(f"total gpu {marks[0][1].elapsed_time(marks[-1][1]):.2f} ms")
