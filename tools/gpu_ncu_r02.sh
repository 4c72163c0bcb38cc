#!/bin/bash
# ncu --set full captures of the three tcgen05 conv flavours (kernel names are demangled with (int) casts)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv3x3_patch_kernel<\(int\)64" -s 8 -c 2 -f -o gpurun_out/prof_patch64_r02 $B > gpurun_out/r02_ncu_patch.log 2>&1
tail -1 gpurun_out/r02_ncu_patch.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel<\(int\)128, \(int\)2, \(int\)64, \(int\)1>" -s 60 -c 6 -f -o gpurun_out/prof_gemm_128_r02 $B > gpurun_out/r02_ncu_gemm128.log 2>&1
tail -1 gpurun_out/r02_ncu_gemm128.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel<\(int\)256, \(int\)2, \(int\)64, \(int\)2>" -s 33 -c 4 -f -o gpurun_out/prof_gemm_pair_r02 $B > gpurun_out/r02_ncu_pair.log 2>&1
tail -1 gpurun_out/r02_ncu_pair.log | cut -c1-120
ls -la gpurun_out/*_r02.ncu-rep
