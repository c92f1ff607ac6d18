"""In-tree build of librestoragen.so (nvcc, sm_100a only).  No torch headers, no pybind: plain C ABI."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
# RG_LIB_OUT=<path>: build an experimental variant (e.g. RG_NVCC_EXTRA=-DRG_GEMM_TUNING) next to the product library
LIB = Path(os.environ["RG_LIB_OUT"]).resolve() if os.environ.get("RG_LIB_OUT") else HERE / "librestoragen.so"
OBJDIR = HERE / ("build_alt" if os.environ.get("RG_LIB_OUT") else "build")
# the fp16 parity build: the same sources with -DRG_OPERAND_F16 (see _lib.py / include/restoragen.h: rg_operand_dtype)
LIB_F16 = HERE / "librestoragen_f16.so"
OBJDIR_F16 = HERE / "build_f16"
SOURCES = ["api.cu", "gemm.cu", "attention.cu", "norm.cu", "elementwise.cu", "metrics.cu", "lpips.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("RG_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; librestoragen.so cannot be built")
    return cand


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [HERE.parent / "include" / "restoragen.h", Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link librestoragen.so (bf16 operands, the product) and
    librestoragen_f16.so (fp16 parity mode) next to this file.  Returns the product library."""
    lib = _build_one(LIB, OBJDIR, [], force, verbose)
    if not os.environ.get("RG_LIB_OUT"):
        _build_one(LIB_F16, OBJDIR_F16, ["-DRG_OPERAND_F16"], force, verbose)
    return lib


def _build_one(LIB: Path, objdir: Path, extra: list, force: bool, verbose: bool) -> Path:
    digest = _digest() + " " + " ".join(NVCC_FLAGS + extra)
    STAMP = objdir / "stamp.txt"
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
