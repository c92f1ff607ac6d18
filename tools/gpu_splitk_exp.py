"""Where does the split-K GEMM's time go?  Runs the 8x8- and 16x16-level convs of config 1 (UNet batch 2) inside a CUDA
graph with parts of epilogue_splitk switched off (RG_GEMM_DEBUG bits 8 = no partial dump, 16 = no fence / counter /
fix-up, 32 = fix-up reads slice 0 only) -- needs the RG_GEMM_TUNING build:
    RG_LIB=tools/micro/librestoragen_tune.so RG_GEMM_DEBUG=<bits> python tools/gpu_splitk_exp.py"""
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
if "nosplit" in sys.argv:
    ops.SPLITK = False
if "nopdl" in sys.argv:
    _lib.load().rg_set_pdl(0)
dev = torch.device("cuda", 0)


def graph_time(fn, reps_in_graph=40, replays=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps_in_graph):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (replays * reps_in_graph)


out = []
for (H, Cin, Cout) in ((8, 1280, 1280), (8, 2560, 1280), (16, 1280, 1280), (16, 2560, 1280), (32, 1280, 640)):
    x = torch.randn((2, H, H, Cin), device=dev).to(torch.bfloat16)
    w = (torch.randn((Cout, 9 * Cin), device=dev) / (9 * Cin) ** 0.5).to(torch.bfloat16)
    b = torch.randn((Cout,), device=dev)
    of = torch.zeros((2, H, H, Cout), device=dev)
    us = graph_time(lambda: ops.conv2d(x, w, kh=3, kw=3, pad_t=1, pad_l=1, bias=b, out_f32=of))
    out.append(f"{H}x{H} {Cin}->{Cout}: {us:6.2f} us")
print(f"dbg={os.environ.get('RG_GEMM_DEBUG', '0'):>3s} argv={sys.argv[1:]}  " + " | ".join(out), flush=True)
