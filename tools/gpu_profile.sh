#!/bin/bash
# ncu evidence for profiles/: launch list of one bench command + full captures of the key kernels.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 250 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 64 -c 6 -f -o gpurun_out/prof_gemm_r01 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:"stem_tc|xcorr_n|maxpool|small_conv" -s 20 -c 16 -f -o gpurun_out/prof_misc_r01 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_misc.log 2>&1
tail -2 gpurun_out/ncu_misc.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
