#!/bin/bash
# ncu metric pass over the PREPROCESS and DECODE kernels (mel front-end, text ConvNeXt-V2, Vocos backbone, fused
# iSTFT / overlap-add) at the bench shape; the DiT loop kernels are filtered out by name.
R=${ROUND:-r02}
mkdir -p gpurun_out
S="python tools/prof_stages.py 1"
$S > gpurun_out/plain_stages.log 2>&1 &&
timeout -s KILL 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,launch__registers_per_thread \
    -k regex:"mel_kernel|rms_scale|istft_ola|dwconv_rows|voc_im2col|grn_|text_gather|cat_cond|ln_kernel<4|philox|noise_to_bf16|gemm_kernel" \
    -c 200 --csv --log-file gpurun_out/${R}_ncu_stages.csv $S > gpurun_out/ncu_stages.log 2>&1
echo "stage metrics rc=$?"
tail -n 2 gpurun_out/plain_stages.log
