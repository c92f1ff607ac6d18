#!/bin/bash
# Sweep of the attention tuning knobs (ab_libs/lib_tune.so = build with -DRG_ATTN_TUNING) inside ONE gpurun call.
# usage: tools/gpu_attn_sweep.sh <out-suffix> "<lib tag>:<issuers>:<skew>" ...
sfx=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k attn > gpurun_out/pytest_attn_$sfx.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_attn_$sfx.log
tail -2 gpurun_out/pytest_attn_$sfx.log
: > gpurun_out/attn_ab_$sfx.txt; : > gpurun_out/attn_trace_$sfx.txt
RG_LIB=ab_libs/lib_base.so timeout 300 python tools/gpu_attn_ab.py base >> gpurun_out/attn_ab_$sfx.txt 2>&1
for cfg in "$@"; do
  IFS=: read lib iss skew <<< "$cfg"
  export RG_ATTN_ISSUERS=$iss RG_ATTN_SKEW=$skew
  RG_LIB=ab_libs/lib_$lib.so timeout 300 python tools/gpu_attn_ab.py "$lib-iss$iss-skew$skew" >> gpurun_out/attn_ab_$sfx.txt 2>&1
  echo "== trace $cfg" >> gpurun_out/attn_trace_$sfx.txt
  RG_LIB=ab_libs/lib_$lib.so timeout 300 python tools/gpu_attn_trace.py 40 4096 >> gpurun_out/attn_trace_$sfx.txt 2>&1
done
unset RG_ATTN_ISSUERS RG_ATTN_SKEW
timeout 300 python tools/gpu_attn_ab.py product >> gpurun_out/attn_ab_$sfx.txt 2>&1
cat gpurun_out/attn_ab_$sfx.txt
grep -E "==|per launch|period|skew" gpurun_out/attn_trace_$sfx.txt
