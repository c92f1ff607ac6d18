"""fp16 parity mode (RESTORAGEN_OPERAND_DTYPE=fp16 -> librestoragen_f16.so: fp16 storage and fp16 tensor-core operands end to end,
the reference's CUDA dtype, src/inference.py:57) against the fp32 oracle: UNet step, VAE encode / decode, one denoise run.
Prints one JSON line.  Run in its own process: the operand dtype is chosen at import.
    RESTORAGEN_OPERAND_DTYPE=fp16 python tools/gpu_fp16_mode_check.py"""
import json
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ.setdefault("RESTORAGEN_OPERAND_DTYPE", "fp16")
import torch
import model_cases as mc
from image_restoration_and_enhancement_b200 import _lib, ops

out = {"library": _lib.LIB_PATH.name, "operand_dtype": str(ops.OPERAND_DTYPE), "rg_operand_dtype": int(_lib.load().rg_operand_dtype())}
out["unet_rel_l2"] = mc.case_unet(in_channels=4, B=1, h=64, w=64, cfg=True, t=501.0)[0]
out["unet9_rel_l2"] = mc.case_unet(in_channels=9, B=1, h=64, w=64, cfg=True, t=562.0)[0]
out["vae_encode_rel_l2"] = mc.case_vae_encode(1)[0]
e, _, p = mc.case_vae_decode(1)
out["vae_decode_rel_l2"], out["vae_decode_psnr"] = e, p
r = mc.case_pipeline("denoise")
out["denoise"] = {"steps": r["steps"], "timesteps_match": r["timesteps_match"], "max_unet_step_rel": max(r["unet_step_rel"]),
                  "final_latents_rel": r["final_latents_rel"], "psnr": r["psnr"]}
print(json.dumps(out))
