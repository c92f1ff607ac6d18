"""LayerNorm timed alone (CUDA events), inputs rotated through buffers larger than the L2.  RG_LIB=<path> for another build."""
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
for rows, C in ((65536, 320), (16384, 640), (4096, 1280), (1024, 1280)):
    nbuf = max(2, min(8, int(400e6 // (rows * C * 4)) + 1))
    xs = [torch.randn((rows, C), device="cuda") for _ in range(nbuf)]
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for x in xs:
        ops.layernorm(x, g, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 * nbuf
    e0.record()
    for i in range(reps):
        ops.layernorm(xs[i % nbuf], g, b)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{_lib.LIB_PATH.name:28s} layernorm rows={rows:6d} C={C:4d}: {us:7.1f} us  {rows * C * 6 / us / 1e3:6.0f} GB/s")
