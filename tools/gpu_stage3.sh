#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --dump-layers gpurun_out/layers_exact.tsv > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
tail -3 gpurun_out/bench_exact.err; cut -c1-400 gpurun_out/bench_exact.json
timeout 900 python bench.py --steps 20 --warmup 3 --precision fast --no-cpu --dump-layers gpurun_out/layers_fast.tsv > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
cut -c1-400 gpurun_out/bench_fast.json
timeout 1200 ncu --set full --clock-control none -k regex:conv_gemm -s 46 -c 30 -o gpurun_out/prof_gemm_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/
