#!/bin/bash
# A/B of the attention kernels on one box: for every VVB200_ATTN variant the attention parity tests, then the
# micro-benchmark at the bench shape (16 sequences x 1501 x 16 heads).  Each step under its own hard timeout.
mkdir -p gpurun_out
: > gpurun_out/attn_ab_summary.txt
for v in ${ATTN_VARIANTS:-1 96 96p 112 112p 80}; do
  VVB200_ATTN=$v timeout -s KILL 240 python -m pytest tests/test_kernels_gpu.py -m gpu -k attention -q -x --no-header \
      -p no:cacheprovider > gpurun_out/attn_ab_test_$v.log 2>&1
  rc=$?
  echo "== variant $v tests rc=$rc: $(tail -n 1 gpurun_out/attn_ab_test_$v.log)" | tee -a gpurun_out/attn_ab_summary.txt
  if [ $rc -ne 0 ]; then tail -n 30 gpurun_out/attn_ab_test_$v.log; continue; fi
  VVB200_ATTN=$v timeout -s KILL 120 python tools/prof_kernels.py attn 20 > gpurun_out/attn_ab_prof_$v.log 2>&1
  echo "   variant $v prof rc=$?: $(grep attention gpurun_out/attn_ab_prof_$v.log)" | tee -a gpurun_out/attn_ab_summary.txt
done
