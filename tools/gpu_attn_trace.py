"""Per-phase cycle breakdown of the attention softmax warps (debug hook rg_debug_attn_trace), d=40, N=4096."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import os
from image_restoration_and_enhancement_b200 import _lib
if os.environ.get('RG_LIB'):
    _lib.LIB_PATH = Path(os.environ['RG_LIB']).resolve()
from image_restoration_and_enhancement_b200 import ops

d = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = _lib.load()
dt = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float16      # the UNet runs the fp16 path
qkv = torch.randn((16, N, 3, 8, d), device="cuda").to(dt)
buf = torch.zeros((18, 64, 8), dtype=torch.int64, device="cuda")
ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
lib.rg_debug_attn_trace.argtypes = [__import__("ctypes").c_void_p]
if len(sys.argv) > 4:                 # which of CTA 0's work items to trace (default: its first)
    lib.rg_debug_attn_trace_item.argtypes = [__import__("ctypes").c_int]
    lib.rg_debug_attn_trace_item(int(sys.argv[4]))
lib.rg_debug_attn_trace(buf.data_ptr())
ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
torch.cuda.synchronize()
lib.rg_debug_attn_trace(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
e1.record(); torch.cuda.synchronize()
print(f"d={d} N={N} {dt}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
t = buf.cpu()
for w in (0, 4, 8, 12):
    tt = t[w]
    n = int((tt[:, 0] > 0).sum())
    if n < 3:
        print("warp", w, "no trace"); continue
    base = int(tt[0, 0])
    print(f"warp {w}: {n} tiles; tile period avg {(int(tt[n-1,0]) - int(tt[1,0])) / (n - 2):.0f} cycles")
    # stamp k -> k+1 durations; stamp 7 -> next tile's stamp 0
    for k in range(8):
        if k < 7:
            dur = (tt[1:n, k + 1] - tt[1:n, k]).float()
        else:
            dur = (tt[2:n, 0] - tt[1:n - 1, 7]).float()
        phase = ["wait s_full", "tcgen05.ld S", "s_free, mask, max, rescale, exp turn", "exponentials", "wait pv_done(G-1)",
                 "tcgen05.st P + wait", "fence + arrive p_full", "loop top"][k]
        print(f"   {phase:26s} avg {dur.mean():8.0f}  min {dur.min():6.0f}  max {dur.max():6.0f}")
    print("   first tiles (start offsets):", [int(tt[i, 0]) - base for i in range(min(n, 8))])
w0, w4 = t[0], t[4]
print("warp4 - warp0 start skew per tile:", [int(w4[i, 0]) - int(w0[i, 0]) for i in range(0, 32, 4)])
print("warp0 exp-phase start vs warp4 exp-phase start:", [int(w4[i, 3]) - int(w0[i, 3]) for i in range(0, 32, 4)])
