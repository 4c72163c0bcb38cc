#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_gemm|stem_tc|xcorr_n|maxpool|crop_k|small_conv|deconv_k|gather_corr|select_k" --csv --log-file gpurun_out/traffic_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-150; wc -l gpurun_out/traffic_r01.csv
