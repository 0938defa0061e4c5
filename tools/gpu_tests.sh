#!/bin/bash
# Runs the GPU parity tests group by group, each under its own hard timeout so a hung kernel cannot
# hold the box; logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/nproc.txt
run() {  # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout -s KILL $to python -m pytest "$@" -q -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "== $name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 12 gpurun_out/$name.log
  return $rc
}
: > gpurun_out/summary.txt
run gemm_plain 240 tests/test_kernels_gpu.py -m gpu -k "gemm_plain or persistent" || exit 0
run gemm_pair 240 tests/test_kernels_gpu.py -m gpu -k "gemm_pair" || exit 0
run gemm_epi 240 tests/test_kernels_gpu.py -m gpu -k "activation or gate_residual or rope"
run conv 180 tests/test_kernels_gpu.py -m gpu -k "conv_rows"
run ln 120 tests/test_kernels_gpu.py -m gpu -k "ln_modulate"
run attn 240 tests/test_kernels_gpu.py -m gpu -k "attention"
run engine 400 tests/test_engine_gpu.py -m gpu
run prompt 300 tests/test_prompt_cache_gpu.py tests/test_crossfade_gpu.py -m gpu
run path 600 tests/test_path_gpu.py tests/test_configs_gpu.py tests/test_scheduler_gpu.py tests/test_edge_gpu.py -m gpu
run parity_full 600 tests/test_parity_full_gpu.py -m gpu -s
