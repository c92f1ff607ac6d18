#!/bin/bash
# Config-5 sweep (100 images x 4 tasks, PSNR/SSIM/LPIPS) at 1..N GPUs of one box; one JSON per world size in gpurun_out/.
#   gpurun --gpus 8 -- bash tools/gpu_sweep_scaling.sh 8 100
MAXG=${1:-8}; IMAGES=${2:-100}
export SWEEP_IMAGES=$IMAGES
for n in 1 2 4 8; do
  [ $n -le $MAXG ] || continue
  if [ $n -eq 1 ]; then
    python -m image_restoration_and_enhancement_b200.sweep > gpurun_out/sweep_r02_${n}gpu.json 2> gpurun_out/sweep_r02_${n}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      -m image_restoration_and_enhancement_b200.sweep > gpurun_out/sweep_r02_${n}gpu.json 2> gpurun_out/sweep_r02_${n}gpu.err
  fi
  echo "== $n GPU(s): rc=$?"; tail -c 400 gpurun_out/sweep_r02_${n}gpu.json; echo
done
python - <<'PY'
import json, glob
runs = {}
for f in sorted(glob.glob("gpurun_out/sweep_r02_*gpu.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    runs[d["_world_size"]] = d
base = runs.get(1)
for w, d in sorted(runs.items()):
    tasks = [k for k in d if not k.startswith("_")]
    same = all(d[t]["metrics"] == base[t]["metrics"] for t in tasks) if base else None
    print(f"world {w}: wall {d['_wall_seconds_incl_model_setup']:.1f} s incl. setup; per task s: "
          + ", ".join(f"{t} {d[t]['seconds_rank0']:.2f}" for t in tasks)
          + f"; sum {sum(d[t]['seconds_rank0'] for t in tasks):.2f} s; statistics bit-identical to 1 GPU: {same}")
PY
