"""SM clock and board power while the 4096-token attention launch runs back to back for ~3 s (nvidia-smi sampled every
100 ms) -- is a variant slower in cycles, or is the chip clocking down under it?   RG_LIB=<path> python tools/gpu_attn_clocks.py [tag]"""
import os, subprocess, sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from image_restoration_and_enhancement_b200 import _lib
if os.environ.get("RG_LIB"):
    _lib.LIB_PATH = Path(os.environ["RG_LIB"]).resolve()
from image_restoration_and_enhancement_b200 import ops
tag = sys.argv[1] if len(sys.argv) > 1 else _lib.LIB_PATH.name
g = torch.Generator(device="cuda").manual_seed(1)
q = (torch.randn((16, 4096, 8, 40), device="cuda", generator=g) * 0.7).half()
k = (torch.randn((16, 4096, 8, 40), device="cuda", generator=g) * 0.7).half()
v = torch.randn((16, 4096, 8, 40), device="cuda", generator=g).half()
rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                         "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
for _ in range(5):
    ops.attention(q, k, v, 40 ** -0.5)
torch.cuda.synchronize()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(50):
        ops.attention(q, k, v, 40 ** -0.5)
    n += 50
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
time.sleep(0.2); proc.terminate()
mid = rows[len(rows) // 3:]
clk = sorted(int(r.split(",")[0]) for r in mid if r.split(",")[0].strip().isdigit())
pw = sorted(float(r.split(",")[1]) for r in mid if r.count(",") >= 1)
print(f"{tag:12s} {e0.elapsed_time(e1) * 1e3 / n:7.1f} us per launch over {n} launches; SM clock median {clk[len(clk) // 2] if clk else None} MHz "
      f"(min {clk[0] if clk else None}), power median {pw[len(pw) // 2] if pw else None} W; power-cap samples {sum(' Active' in r.split(',')[2] for r in mid if r.count(',') >= 2)}/{len(mid)}")
