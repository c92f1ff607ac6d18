"""Top stalled SASS instructions of an .ncu-rep (source page), read here without a GPU.
   python tools/ncu_stalls.py gpurun_out/prof.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
blocks, cur = [], None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = [ln]
        blocks.append(cur)
    elif cur is not None:
        cur.append(ln)
for blk in blocks[:1]:
    name = next(csv.reader([blk[0]]))[1]
    rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    k, src, ex = ci["Warp Stall Sampling (All Samples)"], ci["Source"], ci["Instructions Executed"]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for idx, r in enumerate(rows[1:]):
        try:
            data.append((float(r[k]), idx, r))
        except (ValueError, IndexError):
            pass
    tot = sum(v for v, _, _ in data) or 1.0
    print(name, "total samples", tot)
    agg = {h: 0.0 for h in stall_cols}
    for v, _, r in data:
        for h in stall_cols:
            try:
                agg[h] += float(r[ci[h]])
            except (ValueError, IndexError):
                pass
    print("stall mix:", ", ".join(f"{h[6:]}={100 * a / tot:.1f}%" for h, a in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for v, idx, r in sorted(data, key=lambda x: -x[0])[:top]:
        why = sorted(((float(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"{v:8.0f} {100 * v / tot:5.1f}%  line {idx:5d} exec {r[ex]:>9s}  {r[src].strip()[:90]:90s} {why[0][1]}:{why[0][0]:.0f} {why[1][1]}:{why[1][0]:.0f}")
