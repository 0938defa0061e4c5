#!/bin/bash
# phase timing (clock64 accounting, instrumented library libvvb200_T.so) of the attention kernels
mkdir -p gpurun_out
for v in ${ATTN_VARIANTS:-1 96 96s 112}; do
  echo "=== VVB200_ATTN=$v (instrumented)"
  VVB200_LIB=$PWD/vietvoice-tts_b200/libvvb200_T.so VVB200_ATTN=$v timeout -s KILL 120 python tools/prof_kernels.py attn 10 2>&1 | tee gpurun_out/attn_timing_$v.log | tail -n 25
done
