"""With the K split off (ops.splitk(False), what the sharded sweep uses) an image's result must not depend on the batch it
was sampled in, bit for bit: batch of 5 against single-image calls, img2img and inpaint.   python tools/gpu_batch_invariance.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import model_cases as mc
from image_restoration_and_enhancement_b200 import ops
from image_restoration_and_enhancement_b200.pipelines import StableDiffusionImg2ImgPipeline, StableDiffusionInpaintPipeline

g = torch.Generator().manual_seed(5)
pe, ne = torch.randn((1, 77, 768), generator=g).cuda(), torch.randn((1, 77, 768), generator=g).cuda()
nb = 5
imgs = np.stack([mc.synth_image(31 + i, 256, 256) for i in range(nb)])
mask = np.zeros((nb, 256, 256), dtype=np.uint8); mask[:, 64:128, 80:170] = 255
ok = True
with ops.splitk(False):
    for name, cls, extra in (("img2img", StableDiffusionImg2ImgPipeline, {}), ("inpaint", StableDiffusionInpaintPipeline, {"mask_image": mask})):
        pipe = cls.from_random_init(seed=0, device="cuda").to("cuda")
        kw = dict(prompt_embeds=pe, negative_prompt_embeds=ne, strength=0.6, num_inference_steps=10, guidance_scale=5.0, output_type="np_u8")
        both = pipe(image=imgs, generator=[torch.Generator(device="cuda").manual_seed(42) for _ in range(nb)], **extra, **kw).images
        for i in (0, 2, nb - 1):
            ex = {k: v[i:i + 1] for k, v in extra.items()}
            one = pipe(image=imgs[i:i + 1], generator=torch.Generator(device="cuda").manual_seed(42), **ex, **kw).images
            same = bool((both[i:i + 1] == one).all())
            ok &= same
            print(f"{name}: image {i} in a batch of {nb} == alone: {same}" + ("" if same else f"  (PSNR {mc.psnr_u8(both[i:i+1], one):.2f} dB)"))
print("batch-invariant bit for bit:", ok)
sys.exit(0 if ok else 1)
