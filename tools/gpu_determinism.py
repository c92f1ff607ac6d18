"""Diagnostics: run-to-run and batch-shape reproducibility of the CUDA path."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import model_cases as mc
from image_restoration_and_enhancement_b200 import ops
from image_restoration_and_enhancement_b200.unet import UNetB200
from image_restoration_and_enhancement_b200.weights import random_state_dict, unet_param_shapes

sd = random_state_dict(unet_param_shapes(), 0)
um = UNetB200(sd, device="cuda")
g = torch.Generator().manual_seed(1)
lat = torch.randn((2, 32, 32, 4), generator=g).cuda()
ctx = torch.randn((4, 77, 768), generator=g).cuda()
um.prepare_context(ctx)
ts = torch.full((4,), 400.0, device="cuda")
a = um.forward(lat, ts).clone()
b = um.forward(lat, ts).clone()
print("unet run-to-run rel", mc.rel_l2(a, b), "max abs", float((a - b).abs().max()))
# batch 1 (image 0 only) vs batch 2
um.prepare_context(torch.cat([ctx[0:1], ctx[2:3]]))
c = um.forward(lat[0:1].contiguous(), ts[:2]).clone()
print("unet batch2-vs-batch1 rel (uncond half)", mc.rel_l2(a[0:1], c[0:1]), "(cond half)", mc.rel_l2(a[2:3], c[1:2]))

mc.case_pipeline("denoise", 256, 256)
pipe = mc._cache[("pipe", "img2img", 0)]
pe, ne = torch.randn((1, 77, 768), generator=g).cuda(), torch.randn((1, 77, 768), generator=g).cuda()
imgs = np.stack([mc.synth_image(31 + i, 256, 256) for i in range(2)])
kw = dict(prompt_embeds=pe, negative_prompt_embeds=ne, strength=0.5, num_inference_steps=10, guidance_scale=5.0, output_type="np_u8")
for graph in (True, False):
    pipe.use_cuda_graph = graph
    pipe._graphs.clear()
    t1, t2, t3 = {}, {}, {}
    o1 = pipe(image=imgs[0:1], generator=torch.Generator(device="cuda").manual_seed(42), trace=t1, **kw).images
    o2 = pipe(image=imgs[0:1], generator=torch.Generator(device="cuda").manual_seed(42), trace=t2, **kw).images
    ob = pipe(image=imgs, generator=[torch.Generator(device="cuda").manual_seed(42) for _ in range(2)], trace=t3, **kw).images
    print("graph", graph, "same-call PSNR", mc.psnr_u8(o1, o2), "batched-vs-single PSNR", mc.psnr_u8(ob[0:1], o1))
    print("  init latents rel", mc.rel_l2(t1["init_latents"], t2["init_latents"]), mc.rel_l2(t3["init_latents"][0:1], t1["init_latents"]))
    for i in range(len(t1["latents"])):
        print("  step", i, "eps rel run2run", mc.rel_l2(t1["eps_uc"][i], t2["eps_uc"][i]),
              "latents rel", mc.rel_l2(t1["latents"][i], t2["latents"][i]),
              "| batched eps", mc.rel_l2(t3["eps_uc"][i][0:1], t1["eps_uc"][i][0:1]), "lat", mc.rel_l2(t3["latents"][i][0:1], t1["latents"][i]))
    print("  final latents", mc.rel_l2(t1["final_latents"], t2["final_latents"]), "decoded", mc.rel_l2(t1["decoded"], t2["decoded"]))
