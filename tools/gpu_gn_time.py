"""GroupNorm kernels timed alone (CUDA events), statistics and apply passes separately, inputs rotated through
buffers larger than the L2 so every launch reads HBM.  RG_LIB=<path> loads another build of the library.

    python tools/gpu_gn_time.py [tag]   -> gpurun_out/gn_time_<tag>.txt
"""
import ctypes as C
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
from image_restoration_and_enhancement_b200._lib import RG_DT_BF16, RG_DT_F32, RgGn

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
lib = _lib.load()
dev = torch.device("cuda", 0)
lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


SHAPES = [  # N, HW, C, fp32 input
    (16, 4096, 320, True), (16, 4096, 320, False), (16, 4096, 640, True), (16, 1024, 640, True), (16, 1024, 640, False),
    (16, 1024, 1920, True), (16, 256, 1280, True), (16, 256, 1280, False), (16, 256, 2560, True), (16, 64, 1280, True), (16, 64, 2560, True),
    (8, 262144, 128, False), (8, 65536, 256, False), (8, 16384, 512, False), (8, 4096, 512, False), (1, 4096, 320, True),
    (2, 4096, 320, True), (2, 4096, 320, False), (2, 1024, 640, True), (1, 262144, 128, False), (1, 65536, 256, False)]
say(f"library: {_lib.LIB_PATH.name}")
say(f"{'shape':34s} {'stats us':>9s} {'GB/s':>7s} {'apply us':>9s} {'GB/s':>7s} {'both us':>8s} {'GB/s':>7s}")
for N, HW, Cc, f32in in SHAPES:
    isz = 4 if f32in else 2
    nbytes = N * HW * Cc * isz
    nbuf = max(2, min(8, int(400e6 // nbytes) + 1))
    xs = [torch.randn((N, HW, Cc), device=dev, dtype=torch.float32 if f32in else torch.bfloat16) for _ in range(nbuf)]
    y = torch.empty((N, HW, Cc), device=dev, dtype=torch.bfloat16)
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    ws = ops._gn_workspace(dev, N, 32)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def params(x):
        p = RgGn()
        p.x1, p.C1, p.x2, p.C2 = x.data_ptr(), Cc, None, 0
        p.in_dtype = RG_DT_F32 if f32in else RG_DT_BF16
        p.N, p.HW, p.groups, p.eps = N, HW, 32, 1e-5
        p.gamma, p.beta, p.sums, p.y, p.raw, p.silu = gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), y.data_ptr(), None, 1
        return p
    ps = [params(x) for x in xs]
    reps = 5 * nbuf

    def timed(fn):
        for p in ps:
            fn(p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(ps[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    t_s = timed(lambda p: _lib.check(lib.rg_groupnorm_stats(C.byref(p), st)))
    t_a = timed(lambda p: _lib.check(lib.rg_groupnorm_apply(C.byref(p), st)))
    both = getattr(lib, "rg_groupnorm", None)
    if both is not None:
        t_b = timed(lambda p: _lib.check(lib.rg_groupnorm(C.byref(p), st)))          # one-pass kernel where it applies
    else:
        t_b = timed(lambda p: (_lib.check(lib.rg_groupnorm_stats(C.byref(p), st)), _lib.check(lib.rg_groupnorm_apply(C.byref(p), st))))
    el = N * HW * Cc
    say(f"N={N:2d} HW={HW:6d} C={Cc:4d} {'f32 ' if f32in else 'bf16'} x{nbuf}  {t_s:9.1f} {el * isz / t_s / 1e3:7.0f} {t_a:9.1f} "
        f"{el * (isz + 2) / t_a / 1e3:7.0f} {t_b:8.1f} {el * (2 * isz + 2) / t_b / 1e3:7.0f}")
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"gn_time_{tag}.txt").write_text("\n".join(lines) + "\n")
