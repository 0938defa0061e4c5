#!/bin/bash
# Round profile pass (one gpurun call): ncu launch list of the bench command + `ncu --set full` captures of the
# dominant kernels.  Every ncu run is preceded by the same command run plain (must exit 0).  Outputs in gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu"
$B > gpurun_out/plain_bench.log 2> gpurun_out/plain_bench.err &&
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 420 --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
G="python tools/prof_kernels.py gemm 3"
$G > gpurun_out/plain_gemm.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 2 -c 1 -f \
    -o gpurun_out/gemm_pair_qkv $G > gpurun_out/ncu_gemm_qkv.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 7 -c 1 -f \
    -o gpurun_out/gemm_pair_out $G > gpurun_out/ncu_gemm_out.log 2>&1
echo "gemm captures rc=$?"
A="python tools/prof_kernels.py attn 3"
$A > gpurun_out/plain_attn.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 2 -c 1 -f \
    -o gpurun_out/attn_final $A > gpurun_out/ncu_attn_final.log 2>&1
echo "attention capture rc=$?"
