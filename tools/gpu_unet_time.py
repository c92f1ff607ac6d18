"""UNet evaluation time (CUDA-graph replay) at a few UNet batches.   python tools/gpu_unet_time.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import model_cases as mc
for B, cfg in ((1, False), (1, True), (2, True), (4, False), (8, True)):
    r = mc.time_unet(B=B, cfg=cfg, iters=20)
    print(f"UNet batch {r['Bu']:2d}: {r['ms']:.3f} ms per evaluation ({r['launches']} launches, {r['tflops']:.0f} TFLOP/s)")
