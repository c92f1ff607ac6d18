"""Steady-state cost per launch of the small kernels of one UNet evaluation at UNet batch 2 (config 1): each op is
captured 40x back to back in ONE CUDA graph (a dependent chain, like the real evaluation) and replayed -- no event or
host overhead in the figure, warm instruction / descriptor caches.  This is the number the batch-1 UNet evaluation is
made of (396 launches), unlike per-op CUDA events in eager mode (>= 10 us of launch gap each) or ncu (cold caches).

    python tools/gpu_kernel_floor.py [tag]      -> gpurun_out/floor_<tag>.txt
"""
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
from image_restoration_and_enhancement_b200._lib import RG_ACT_GEGLU

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
if "nosplit" in sys.argv[2:]:
    ops.SPLITK = False
if "nopdl" in sys.argv[2:]:
    _lib.load().rg_set_pdl(0)
if "pdl1" in sys.argv[2:]:
    _lib.load().rg_set_pdl(1)
dev = torch.device("cuda", 0)
bf16, f32, f16 = torch.bfloat16, torch.float32, torch.float16
lines = []


def say(s=""):
    print(s, flush=True)
    lines.append(s)


def graph_time(fn, reps_in_graph=40, replays=20):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps_in_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * reps_in_graph)


def rnd(*shape, dtype=bf16, scale=1.0):
    return (torch.randn(shape, device=dev) * scale).to(dtype)


def gemm_case(name, N, H, W, Cin, Cout, k=1, res=False, f32_out=False, act=0, images_linear=False, x2c=0):
    x = rnd(N, H, W, Cin)
    ktot = k * k * Cin + x2c
    w = rnd(Cout, ktot, scale=ktot ** -0.5)
    b = rnd(Cout, dtype=f32)
    x2 = rnd(N, H, W, x2c) if x2c else None
    Cw = Cout // 2 if act == RG_ACT_GEGLU else Cout
    of = torch.zeros((N, H, W, Cw), dtype=f32, device=dev) if f32_out else None
    ob = None if f32_out else torch.zeros((N, H, W, Cw), dtype=bf16, device=dev)
    r = rnd(N, H, W, Cw, dtype=f32) if res else None

    def fn():
        ops.conv2d(x, w, kh=k, kw=k, pad_t=k // 2, pad_l=k // 2, x2=x2, bias=b, res=r, out_bf16=ob, out_f32=of, act=act)
    us = graph_time(fn)
    fl = 2.0 * N * H * W * Cout * ktot
    say(f"{us:8.2f} us  {fl / us / 1e6:8.1f} TF/s  gemm  {name}: M={N * H * W} N={Cout} K={ktot}")


with torch.no_grad():
    say(f"# per-launch cost inside a CUDA graph (40 dependent launches per graph), splitk={'on' if ops.SPLITK else 'off'} argv={sys.argv[2:]}")
    # transformer linears at UNet batch 2 (tokens as [N,1,HW,C])
    gemm_case("64^2 to_q / proj (K=320)", 2, 1, 4096, 320, 320)
    gemm_case("64^2 out-proj res f32", 2, 1, 4096, 320, 320, res=True, f32_out=True)
    gemm_case("64^2 qkv", 2, 1, 4096, 320, 960)
    gemm_case("64^2 geglu", 2, 1, 4096, 320, 2560, act=RG_ACT_GEGLU)
    gemm_case("64^2 ff-out", 2, 1, 4096, 1280, 320, res=True)
    gemm_case("32^2 out-proj", 2, 1, 1024, 640, 640, res=True, f32_out=True)
    gemm_case("32^2 geglu", 2, 1, 1024, 640, 5120, act=RG_ACT_GEGLU)
    gemm_case("32^2 ff-out", 2, 1, 1024, 2560, 640, res=True)
    gemm_case("16^2 out-proj", 2, 1, 256, 1280, 1280, res=True, f32_out=True)
    gemm_case("16^2 qkv", 2, 1, 256, 1280, 3840)
    gemm_case("16^2 geglu", 2, 1, 256, 1280, 10240, act=RG_ACT_GEGLU)
    gemm_case("16^2 ff-out", 2, 1, 256, 5120, 1280, res=True)
    gemm_case("8^2 out-proj", 2, 1, 64, 1280, 1280, res=True, f32_out=True)
    gemm_case("8^2 ff-out", 2, 1, 64, 5120, 1280, res=True)
    # 3x3 convs
    gemm_case("64^2 conv 320->320", 2, 64, 64, 320, 320, k=3)
    gemm_case("64^2 conv 640->320 + shortcut", 2, 64, 64, 320, 320, k=3, x2c=640, f32_out=True)
    gemm_case("32^2 conv 640->640", 2, 32, 32, 640, 640, k=3)
    gemm_case("32^2 conv 1280->640", 2, 32, 32, 1280, 640, k=3)
    gemm_case("16^2 conv 1280->1280", 2, 16, 16, 1280, 1280, k=3)
    gemm_case("16^2 conv 2560->1280", 2, 16, 16, 2560, 1280, k=3)
    gemm_case("8^2 conv 1280->1280 res", 2, 8, 8, 1280, 1280, k=3, res=True, f32_out=True)
    gemm_case("8^2 conv 2560->1280", 2, 8, 8, 2560, 1280, k=3, f32_out=True)
    # norms
    for (HW, C, in_f32) in ((4096, 320, True), (4096, 320, False), (1024, 640, True), (256, 1280, True), (64, 1280, True)):
        xx = rnd(2, int(HW ** 0.5), int(HW ** 0.5), C, dtype=f32 if in_f32 else bf16)
        ga, be = rnd(C, dtype=f32), rnd(C, dtype=f32)
        us = graph_time(lambda: ops.groupnorm(xx, ga, be, silu=True))
        say(f"{us:8.2f} us  groupnorm N=2 HW={HW} C={C} in={'f32' if in_f32 else 'bf16'}")
    for (rows, C) in ((8192, 320), (2048, 640), (512, 1280)):
        xx = rnd(rows, C, dtype=f32)
        ga, be = rnd(C, dtype=f32), rnd(C, dtype=f32)
        us = graph_time(lambda: ops.layernorm(xx, ga, be))
        say(f"{us:8.2f} us  layernorm rows={rows} C={C}")
    # attention
    for (Nq, Nk, d) in ((4096, 4096, 40), (4096, 77, 40), (1024, 1024, 80), (1024, 77, 80), (256, 256, 160), (256, 77, 160),
                        (64, 64, 160), (64, 77, 160)):
        q = rnd(2, Nq, 8, d, dtype=f16)
        kk = rnd(2, Nk, 8, d, dtype=f16)
        vv = rnd(2, Nk, 8, d, dtype=f16)
        o = torch.empty((2, Nq, 8, d), dtype=bf16, device=dev)
        us = graph_time(lambda: ops.attention(q, kk, vv, d ** -0.5, out=o), reps_in_graph=20)
        say(f"{us:8.2f} us  {4.0 * 2 * 8 * Nq * Nk * d / us / 1e6:8.1f} TF/s  attention B=2 H=8 Nq={Nq} Nk={Nk} d={d}")

    # the whole UNet evaluation (batch 2, config 1) as one graph
    from image_restoration_and_enhancement_b200.unet import UNetB200
    from image_restoration_and_enhancement_b200.weights import random_state_dict, unet_param_shapes
    um = UNetB200(random_state_dict(unet_param_shapes(4), 0), in_channels=4, device="cuda")
    um.prepare_context(rnd(2, 77, 768, dtype=f32))
    lat = rnd(1, 64, 64, 4, dtype=f32)
    tsv = torch.full((2,), 500.0, device=dev)
    us = graph_time(lambda: um.forward(lat, tsv), reps_in_graph=1, replays=20)
    say(f"{us / 1e3:8.3f} ms  UNet evaluation, batch 2, graph replay")

(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"floor_{tag}.txt").write_text("\n".join(lines) + "\n")
