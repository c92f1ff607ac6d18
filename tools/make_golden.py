"""Generate the committed fixtures that pin the oracle / host logic (run once in the build container, where
/root/reference is mounted; the GPU box has no reference tree).

  tests/golden/prompt_ids.json                         token ids of the reference's default prompts
  image_restoration_and_enhancement_b200/data/default_prompt_ids.json   (same content, shipped with the product so
                                                        the default prompts need no tokenizer files at run time)
  tests/golden/scheduler_tables.json                   timestep lists + alpha-bar endpoints
  tests/golden/demo_preprocess.json                    VaeImageProcessor sizes / checksums of data/demo images
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))


def main():
    from transformers import CLIPTokenizer
    tok = CLIPTokenizer.from_pretrained(str(REF / "outputs/models/denoising/best/tokenizer"))
    prompts = [
        "clean high quality photo, no noise, sharp details",                       # src/inference.py:87
        "high quality, detailed, sharp",                                           # :88
        "vibrant realistic natural colors, colorful, high quality photo, detailed, full color, rich colors",  # :89
        "high quality detailed photo",                                             # :90
        "high quality detailed photo, realistic",                                  # app.py:272
        "a photograph, high quality, detailed, sharp",                             # scripts/train_denoising.py:399 (validation)
        "",                                                                        # CFG negative prompt
    ]
    ids = {p: tok(p, padding="max_length", max_length=77, truncation=True)["input_ids"] for p in prompts}
    for dst in (ROOT / "tests/golden/prompt_ids.json",
                ROOT / "image_restoration_and_enhancement_b200/data/default_prompt_ids.json"):
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_text(json.dumps(ids, indent=0))

    from oracle.schedulers import PNDMScheduler, DDIMScheduler, get_timesteps
    tables = {}
    for name, cls, n, strength in (("denoise_pndm_20_0.5", PNDMScheduler, 20, 0.5),
                                   ("colorize_pndm_30_0.75", PNDMScheduler, 30, 0.75),
                                   ("sr_pndm_20_0.8", PNDMScheduler, 20, 0.8),
                                   ("sr_pndm_50_0.8", PNDMScheduler, 50, 0.8),
                                   ("inpaint_ddim_30_0.6", DDIMScheduler, 30, 0.6)):
        s = cls()
        s.set_timesteps(n)
        ts, _ = get_timesteps(s, n, strength)
        tables[name] = {"full": [int(t) for t in s.timesteps], "sliced": [int(t) for t in ts]}
    s = PNDMScheduler()
    tables["alphas_cumprod"] = {"0": float(s.alphas_cumprod[0]), "500": float(s.alphas_cumprod[500]),
                                "999": float(s.alphas_cumprod[999])}
    (ROOT / "tests/golden/scheduler_tables.json").write_text(json.dumps(tables, indent=0))

    from PIL import Image
    from oracle.pipelines import preprocess_image
    demo = {}
    for p in sorted((REF / "data/demo/images").glob("*")):
        im = Image.open(p).convert("RGB")
        t = preprocess_image(im)
        u8 = ((t[0].permute(1, 2, 0).numpy() + 1.0) * 127.5).round().astype(np.uint8)
        demo[p.name] = {"in_size": list(im.size), "out_hw": list(t.shape[2:]),
                        "sha256_u8": hashlib.sha256(u8.tobytes()).hexdigest(),
                        "mean": float(t.mean()), "std": float(t.std())}
    (ROOT / "tests/golden/demo_preprocess.json").write_text(json.dumps(demo, indent=0))
    print("wrote fixtures:", len(ids), "prompts,", len(tables), "tables,", len(demo), "demo images")


if __name__ == "__main__":
    main()
