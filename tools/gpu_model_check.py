"""Model-level parity + timing on the GPU box; writes gpurun_out/model_check.json.
   python tools/gpu_model_check.py [unet] [vae] [time] [pipe_denoise] [pipe_colorize] [pipe_inpaint] [pipe_sr]"""
import json
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import model_cases as mc
    what = sys.argv[1:] or ["unet", "vae", "time", "pipe_denoise"]
    out = {}

    def run(name, fn):
        t0 = time.time()
        try:
            out[name] = fn()
        except Exception as e:
            out[name] = {"error": repr(e), "tb": traceback.format_exc()[-1500:]}
        out[name + "_s"] = round(time.time() - t0, 2)
        print(name, json.dumps(out[name])[:600], flush=True)

    if "unet" in what:
        run("unet_b1_cfg", lambda: mc.case_unet(4, 1, 64, 64, True))
        run("unet_b2_nocfg_t1", lambda: mc.case_unet(4, 2, 64, 64, False, t=1.0))
        run("unet_small_ragged", lambda: mc.case_unet(4, 1, 41, 62, True, t=727.0))
        run("unet9_b1_cfg", lambda: mc.case_unet(9, 1, 64, 64, True))
    if "vae" in what:
        run("vae_encode", lambda: mc.case_vae_encode(1))
        run("vae_decode", lambda: mc.case_vae_decode(1))
        run("vae_decode_ragged", lambda: mc.case_vae_decode(1, 41, 62))
    if "time" in what:
        mc._cache.clear()
        torch.cuda.empty_cache()
        run("time_unet_b1", lambda: mc.time_unet(1, True))
        run("time_unet_b8", lambda: mc.time_unet(8, True))
        run("time_unet_b8_eager", lambda: mc.time_unet(8, True, graph=False))
    for task in ("denoise", "colorize", "sr", "inpaint"):
        if "pipe_" + task in what:
            run("pipe_" + task, lambda: mc.case_pipeline(task))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "model_check.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
