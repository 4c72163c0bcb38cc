#!/bin/bash
# ncu --set full of the non-GEMM kernels of one step on the final tree (stem, max-pool, engine xcorr, selection, refine tail)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"stem_tc_kernel|maxpool_kernel|xcorr_nhwc_kernel|select_kernel|deconv_kernel|small_conv3x3|crop_kernel|gather_corr_kernel" \
  -s 60 -c 48 -f -o gpurun_out/prof_misc_r02 $B > gpurun_out/r02_ncu_misc.log 2>&1
tail -1 gpurun_out/r02_ncu_misc.log | cut -c1-120
ls -la gpurun_out/prof_misc_r02.ncu-rep
