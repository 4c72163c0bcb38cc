#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider -k "two_lane or sharp_b1 or batched" 2>&1 | tail -3
timeout 900 python bench.py --steps 50 --warmup 3 --dump-layers gpurun_out/layers_exact.tsv > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
tail -3 gpurun_out/bench_exact.err; cut -c1-200 gpurun_out/bench_exact.json
timeout 900 python bench.py --steps 50 --warmup 3 --precision fast --no-cpu --dump-layers gpurun_out/layers_fast.tsv > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
cut -c1-200 gpurun_out/bench_fast.json
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"256, .int.2, .int.64, .int.2" -s 33 -c 4 -f -o gpurun_out/prof_gemm_pair_r01 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"128, .int.2, .int.64, .int.1" -s 60 -c 6 -f -o gpurun_out/prof_gemm_128_r01 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm128.log 2>&1
tail -2 gpurun_out/ncu_gemm128.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
