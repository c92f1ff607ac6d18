#!/bin/bash
# A/B of attention-kernel builds inside ONE gpurun call: "product" = the in-tree library, any other tag = ab_libs/lib_<tag>.so
# (built with RG_LIB_OUT / RG_NVCC_EXTRA).   usage: tools/gpu_attn_variants.sh <out-suffix> <tag> [<tag> ...]
sfx=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k attn > gpurun_out/pytest_attn_$sfx.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_attn_$sfx.log
tail -3 gpurun_out/pytest_attn_$sfx.log
: > gpurun_out/attn_ab_$sfx.txt
for rep in 1 2; do
  for tag in "$@"; do
    if [ "$tag" = product ]; then lib=; else lib=ab_libs/lib_$tag.so; fi
    RG_LIB=$lib timeout 300 python tools/gpu_attn_ab.py $tag >> gpurun_out/attn_ab_$sfx.txt 2>&1
  done
done
cat gpurun_out/attn_ab_$sfx.txt
