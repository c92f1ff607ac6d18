"""Config-5 sweep slice on one GPU under the three hand-off modes and both metric back ends: the metrics must be
identical (float64 bits) for disk vs memory and for cpu vs gpu scoring; prints the wall time of each combination.

    python tools/gpu_sweep_modes.py [n_images]   -> gpurun_out/sweep_modes.txt
"""
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from image_restoration_and_enhancement_b200 import sweep
from image_restoration_and_enhancement_b200.inference import RestorationPipeline

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


cfg = {t: {"fine_tuned_dir": "nonexistent", "pretrained_id": "", "random_init": 1000 if t == "inpaint" else 0} for t in sweep.TASKS}
pipe = RestorationPipeline(device="cuda", config=cfg, seed=42, strict=True)
for task in ("denoise", "colorize", "inpaint", "sr"):
    sweep.run_task(pipe, task, 8, handoff_mode="none", metrics_backend="gpu")          # warm-up: weights, graphs
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mode, backend, ov in (("memory", "gpu", True), ("memory", "gpu", False), ("memory", "cpu", False), ("disk", "cpu", False),
                                  ("none", "gpu", False)):
            _, vals, secs = sweep.run_task(pipe, task, n, handoff_mode=mode, metrics_backend=backend, workdir=tmp, overlap=ov)
            res[(mode, backend)] = vals if (mode, backend) not in res else res[(mode, backend)]
            assert res[(mode, backend)] == vals
            say(f"{task:9s} handoff={mode:6s} metrics={backend} overlap={int(ov)}: {secs:7.2f} s for {n} images = {n / secs:6.2f} img/s   "
                f"psnr[0]={vals['psnr'][0]!r} ssim[0]={vals['ssim'][0]!r}")
    same_backend = res[("memory", "gpu")] == res[("memory", "cpu")]
    same_handoff = res[("memory", "cpu")] == res[("disk", "cpu")]
    say(f"{task:9s} gpu metrics == cpu metrics (bitwise): {same_backend};  memory hand-off == disk hand-off (bitwise): {same_handoff}")
    assert same_backend and same_handoff
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "sweep_modes.txt").write_text("\n".join(lines) + "\n")
