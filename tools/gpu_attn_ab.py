"""Attention launch times for the UNet's shapes (fp16 operands, batch 16 x 8 heads), CUDA-event timed.
RG_LIB=<path> loads another build of the library, so variants can be compared inside ONE gpurun call (box-to-box clock
differences are several percent).   python tools/gpu_attn_ab.py [tag]"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import _lib
if os.environ.get("RG_LIB"):
    _lib.LIB_PATH = Path(os.environ["RG_LIB"]).resolve()
from image_restoration_and_enhancement_b200 import ops

tag = sys.argv[1] if len(sys.argv) > 1 else _lib.LIB_PATH.name
SHAPES = [(40, 4096, 4096), (80, 1024, 1024), (160, 256, 256), (160, 64, 64), (40, 4096, 77), (80, 1024, 77), (160, 256, 77), (160, 64, 77)]
out = []
for d, nq, nk in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn((16, nq, 8, d), device="cuda", generator=g) * 0.7).half()
    k = (torch.randn((16, nk, 8, d), device="cuda", generator=g) * 0.7).half()
    v = torch.randn((16, nk, 8, d), device="cuda", generator=g).half()
    for _ in range(3):
        ops.attention(q, k, v, d ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ops.attention(q, k, v, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) * 1e3 / reps)
# weights: launches per UNet evaluation (5, 5, 5, 1 self; 5, 5, 5, 1 cross)
w = [5, 5, 5, 1, 5, 5, 5, 1]
print(f"{tag:28s} " + " ".join(f"{t:7.1f}" for t in out) + f"   | per UNet eval {sum(a * b for a, b in zip(out, w)) / 1e3:6.3f} ms")
