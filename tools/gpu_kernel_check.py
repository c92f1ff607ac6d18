"""Run every kernel parity case of tests/kernel_cases.py in its own process (a faulting kernel poisons the
CUDA context) and write a table to gpurun_out/kernel_check.txt.   python tools/gpu_kernel_check.py [pattern]"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def run_one(name: str) -> None:
    import kernel_cases as kc
    err, tol = kc.CASES[name]()
    print(json.dumps({"case": name, "err": err, "tol": tol, "ok": bool(err <= tol)}))


def main() -> int:
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(sys.argv[2])
        return 0
    import kernel_cases as kc
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    lines, bad = [], 0
    for name in kc.CASES:
        if pat and pat not in name:
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--one", name], capture_output=True, text=True, timeout=180)
            last = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
            if r.returncode == 0 and last:
                d = json.loads(last[-1])
                status = "ok  " if d["ok"] else "FAIL"
                bad += 0 if d["ok"] else 1
                line = f"{status} {name:28s} err={d['err']:.3e} tol={d['tol']:.1e} ({time.time() - t0:.1f}s)"
            else:
                bad += 1
                tail = (r.stderr.strip().splitlines() or ["?"])[-1][:300]
                line = f"ERR  {name:28s} rc={r.returncode} {tail}"
        except subprocess.TimeoutExpired:
            bad += 1
            line = f"HANG {name:28s} (timeout)"
        print(line, flush=True)
        lines.append(line)
    (out / "kernel_check.txt").write_text("\n".join(lines) + f"\nfailed: {bad}\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
