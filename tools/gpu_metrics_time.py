"""Times the GPU metric pass (csrc/metrics.cu) against the CPU bookkeeping on one batch of 8 512x512 images and checks
bit equality.   python tools/gpu_metrics_time.py  -> gpurun_out/metrics_time.txt"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

from image_restoration_and_enhancement_b200 import metrics, synth

lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


B = 8
data = synth.batch("denoise", range(B))
gt, pred = data["gt"], data["input"]
dgt, dpred = torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda()
calc = metrics.MetricsCalculator(use_lpips=False)
t0 = time.perf_counter()
cpu = [(calc.calculate_psnr(pred[i], gt[i]), calc.calculate_ssim(pred[i], gt[i])) for i in range(B)]
t_cpu = (time.perf_counter() - t0) / B
for _ in range(3):
    p, s = metrics.psnr_ssim_device(dpred, dgt)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 10
for _ in range(reps):
    p, s = metrics.psnr_ssim_device(dpred, dgt)
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / reps / B
same = all(np.float64(a).tobytes() == np.float64(b).tobytes() for (a, c), b, d in zip(cpu, p, s) for a, b in ((a, b), (c, d)))
say(f"batch of {B} 512x512x3 u8 images (synthetic denoise pairs)")
say(f"CPU (numpy/scipy float64, 1 thread): {t_cpu * 1e3:8.2f} ms / image")
say(f"GPU (librestoragen, incl. D2H of the partial sums and the host tail): {t_gpu * 1e3:8.3f} ms / image   ({t_cpu / t_gpu:.0f}x)")
say(f"bit-identical PSNR and SSIM for all {B} images: {same}")
say(f"example: psnr {p[0]!r} ssim {s[0]!r}")
# device-only kernel time
import ctypes as C
from image_restoration_and_enhancement_b200 import _lib
lib = _lib.load()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N, H, W, Cc = dpred.shape
sse = torch.empty((N,), dtype=torch.int64, device="cuda")
smap = torch.empty((N, Cc, H - 6, W - 6), dtype=torch.float64, device="cuda")
csum = torch.empty((N, Cc, lib.rg_metrics_ssim_chunks(H, W)), dtype=torch.float64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, fn in (("sse", lambda: lib.rg_metrics_sse_u8(dpred.data_ptr(), dgt.data_ptr(), N, H * W * Cc, sse.data_ptr(), st)),
                 ("ssim", lambda: lib.rg_metrics_ssim_u8(dpred.data_ptr(), dgt.data_ptr(), N, H, W, Cc, 6.5025, 58.5225, 49 / 48,
                                                         smap.data_ptr(), csum.data_ptr(), st))):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    say(f"kernel time {name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per batch of {B}")
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "metrics_time.txt").write_text("\n".join(lines) + "\n")
