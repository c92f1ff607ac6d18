"""Times a few GEMM shapes back to back (CUDA events, 20 reps each).  python tools/gpu_gemm_time.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import ops
from image_restoration_and_enhancement_b200._lib import RG_ACT_GEGLU

torch.manual_seed(0)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

M = 65536
x320 = torch.randn((M, 320), device="cuda").to(torch.bfloat16)
x1280 = torch.randn((M, 1280), device="cuda").to(torch.bfloat16)
w = lambda n, k: (torch.randn((n, k), device="cuda") / k ** 0.5).to(torch.bfloat16)
w320, w960, w2560, wff = w(320, 320), w(960, 320), w(2560, 320), w(320, 1280)
b320 = torch.randn((320,), device="cuda"); b2560 = torch.randn((2560,), device="cuda")
r = torch.randn((M, 320), device="cuda")
x4 = torch.randn((16, 64, 64, 320), device="cuda").to(torch.bfloat16)
w3 = w(320, 2880)
cases = {
    "lin N=320 K=320 bf16out": lambda: ops.linear(x320, w320, bias=b320, out_bf16=True),
    "lin N=320 K=320 f32out": lambda: ops.linear(x320, w320, bias=b320, out_f32=True),
    "lin N=320 K=320 res f32 in/out": lambda: ops.linear(x320, w320, bias=b320, res=r, out_f32=r.view(1, 1, M, 320)),
    "lin N=960 K=320 bf16out": lambda: ops.linear(x320, w960, out_bf16=True),
    "geglu N=2560 K=320": lambda: ops.linear(x320, w2560, bias=b2560, act=RG_ACT_GEGLU, out_bf16=True),
    "lin N=320 K=1280 res f32 -> bf16": lambda: ops.linear(x1280, wff, bias=b320, res=r, out_bf16=True),
    "conv3x3 320->320 64x64x16": lambda: ops.conv2d(x4, w3, kh=3, kw=3, pad_t=1, pad_l=1, bias=b320, out_bf16=True),
}
for name, fn in cases.items():
    print(f"{name:36s} {t(fn):8.1f} us", flush=True)
