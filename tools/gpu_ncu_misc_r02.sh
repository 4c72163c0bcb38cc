#!/bin/bash
# ncu --set full of the non-GEMM kernels of one step on the final tree (stem, max-pool, engine xcorr, selection, refine
# tail); the report is condensed on the box (the raw .ncu-rep of 48 launches exceeds gpurun's 64 MiB return limit)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled \
  -k regex:"stem_tc_kernel|maxpool_kernel|xcorr_nhwc_kernel|select_kernel|deconv_kernel|small_conv3x3|crop_kernel|gather_corr_kernel" \
  -s 60 -c 24 -f -o /tmp/prof_misc_r02 $B > gpurun_out/r02_ncu_misc.log 2>&1
tail -1 gpurun_out/r02_ncu_misc.log | cut -c1-120
python tools/ncu_summary.py /tmp/prof_misc_r02.ncu-rep gpurun_out/r02_misc_kernels_ncu_full.csv
wc -l gpurun_out/r02_misc_kernels_ncu_full.csv
