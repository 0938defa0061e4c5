#!/bin/bash
# back-to-back A/B of environment-switchable variants on one box: tools/bench_ab.sh "VAR=val VAR2=val" "..." ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout -s KILL 300 python bench.py --no-cpu --no-configs --steps ${AB_STEPS:-5} --warmup 3 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  rc=$?
  python - "$cfg" $rc gpurun_out/ab_$i.json <<'PY'
import json, sys
cfg, rc, path = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.load(open(path)); k = d["kernels"]
    print(f"{cfg:40s} rc={rc} value {d['value']:.2f}  step {k['graph_step_ms']:.3f} ms  qkv {k['gemm_qkv_ms']:.2f} out {k['gemm_out_ms']:.2f} "
          f"ff1 {k['gemm_ff1_ms']:.2f} ff2 {k['gemm_ff2_ms']:.2f} attn {k['attention_ms']:.2f} ln {k['ln_mod_ms']:.2f}  p50 {d['latency']['p50_ms']:.1f}  sm {d['clocks']['sm_mhz']}")
except Exception as ex:
    print(f"{cfg:40s} rc={rc} FAILED {ex}")
PY
done | tee -a gpurun_out/ab_summary.txt
