#!/bin/bash
# Round profile pass (one gpurun call): ncu launch list of the bench command + `ncu --set full` captures of the
# dominant kernels + a metric pass over the preprocess / decode kernels.  Every ncu run is preceded by the same command
# run plain (must exit 0).  Reports are condensed ON THE BOX (tools/ncu_summary.py) and deleted: gpurun_out/ is capped
# at 64 MiB.  Outputs: gpurun_out/<round>_launches.csv, <round>_ncu_full_summaries.txt, <round>_ncu_stages.csv
R=${ROUND:-r02}
mkdir -p gpurun_out
SUM=gpurun_out/${R}_ncu_full_summaries.txt
: > $SUM
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-configs"
$B > gpurun_out/plain_bench.log 2> gpurun_out/plain_bench.err &&
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 420 --csv \
    --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
full() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout -s KILL 300 ncu --set full --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" \
      > gpurun_out/ncu_$name.log 2>&1 && python tools/ncu_summary.py /tmp/$name.ncu-rep >> $SUM
  echo "$name rc=$?"
  rm -f /tmp/$name.ncu-rep
}
G="python tools/prof_kernels.py gemm 3"
$G > gpurun_out/plain_gemm.log 2>&1 && {
  full gemm_pair_qkv gemm_pair 2 $G      # launch order in prof_kernels: qkv bn=256 x5?, bn=512 ... (same indices as round 1)
  full gemm_pair_out gemm_pair 7 $G
  full gemm_pair_ff1 gemm_pair 12 $G
  full gemm_pair_ff2 gemm_pair 17 $G
}
A="python tools/prof_kernels.py attn 3"
$A > gpurun_out/plain_attn.log 2>&1 && full attn attn_kernel 2 $A
S="python tools/prof_stages.py 1"
$S > gpurun_out/plain_stages.log 2>&1 &&
timeout -s KILL 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,launch__registers_per_thread \
    -k regex:"mel_kernel|rms_scale|istft_ola|dwconv_rows|voc_im2col|grn_|text_gather|cat_cond|ln_kernel|philox|noise_to_bf16|gemm_kernel|cfg_euler|conv_pos" \
    -c 120 --csv --log-file gpurun_out/${R}_ncu_stages.csv $S > gpurun_out/ncu_stages.log 2>&1
echo "stage metrics rc=$?"
tail -n 2 gpurun_out/plain_stages.log
du -sh gpurun_out
