"""What ONE rank of the 8-GPU config-5 sweep spends its time on (13 of 100 images per task), measured on one GPU:
rank 0 of an emulated world of 8, per task the host stages (prepare, score) and the sampling calls, in call order.
    python tools/gpu_sweep_rank_profile.py [world=8] [images=100]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import ops, sweep
from image_restoration_and_enhancement_b200.inference import RestorationPipeline
from image_restoration_and_enhancement_b200.lpips import LPIPSB200, random_lpips_state_dict

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_images = int(sys.argv[2]) if len(sys.argv) > 2 else 100
t00 = time.time()
cfg = {t: {"fine_tuned_dir": "nonexistent", "pretrained_id": "", "random_init": (1000 if t == "inpaint" else 0)} for t in sweep.TASKS}
pipe = RestorationPipeline(device="cuda", config=cfg, seed=42, strict=True)
lp = LPIPSB200(random_lpips_state_dict(0), device="cuda:0")
print(f"construct RestorationPipeline + LPIPS: {time.time() - t00:.2f} s")
orig = pipe.process_batch
log = []


def timed_process_batch(ims, task, **kw):
    torch.cuda.synchronize(); t0 = time.time()
    out = orig(ims, task, **kw)
    torch.cuda.synchronize()
    log.append((task, len(ims), time.time() - t0))
    return out


pipe.process_batch = timed_process_batch
with ops.splitk(False):
    for rep in (0, 1):
        for task in sweep.TASKS:
            log.clear()
            t0 = time.time()
            idx, vals, secs = sweep.run_task(pipe, task, n_images, 0, world, lpips_model=lp)
            print(f"pass {rep} {task:9s} {len(idx)} images: task {secs:.2f} s; sampling calls " +
                  ", ".join(f"B={b}: {s:.2f} s" for _, b, s in log) + f"; host prepare/score outside them {secs - sum(s for *_, s in log):.2f} s")
print(f"total {time.time() - t00:.2f} s")
