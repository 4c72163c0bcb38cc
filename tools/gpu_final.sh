#!/bin/bash
# Round-end validation: full GPU test suite, smoke, bench (both precisions), ncu traffic/launch list, ncu full captures.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log 2>&1
cat gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --dump-layers gpurun_out/layers_exact.tsv > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
tail -3 gpurun_out/bench_exact.err; cut -c1-600 gpurun_out/bench_exact.json
timeout 900 python bench.py --steps 20 --warmup 3 --precision fast --no-cpu --dump-layers gpurun_out/layers_fast.tsv > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
tail -3 gpurun_out/bench_fast.err; cut -c1-300 gpurun_out/bench_fast.json
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_traffic.log 2>&1
tail -1 gpurun_out/ncu_traffic.log | cut -c1-150; wc -l gpurun_out/traffic_r01.csv
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel<256, 2, 64, 2>" -s 33 -c 4 -f -o gpurun_out/prof_gemm_pair_r01 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel<128, 2, 64, 1>" -s 60 -c 6 -f -o gpurun_out/prof_gemm_128_r01 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm128.log 2>&1
tail -2 gpurun_out/ncu_gemm128.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
