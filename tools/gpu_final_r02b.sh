#!/bin/bash
# r02 final records on the frozen tree (the full GPU suite, smoke and the latency table of the same tree: gpu_r2m.sh):
# bench (exact + layers / fast / --config 3 / --config 5), ncu launch list with DRAM traffic, ncu --set full of the three
# GEMM flavours, memcheck over the kernels touched last.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python bench.py --steps 20 --warmup 3 --dump-layers gpurun_out/r02_layers_exact.tsv > gpurun_out/r02_bench_exact_n1.json 2> gpurun_out/r02_bench_exact.err
tail -3 gpurun_out/r02_bench_exact.err; cut -c1-300 gpurun_out/r02_bench_exact_n1.json
timeout 900 python bench.py --steps 20 --warmup 3 --precision fast --no-cpu --no-context --dump-layers gpurun_out/r02_layers_fast.tsv > gpurun_out/r02_bench_fast_n1.json 2> gpurun_out/r02_bench_fast.err
cut -c1-200 gpurun_out/r02_bench_fast_n1.json
timeout 900 python bench.py --config 3 --steps 20 --warmup 3 --no-cpu --no-context > gpurun_out/r02_bench_cfg3_rpn_b256.json 2> gpurun_out/r02_cfg3.err
tail -2 gpurun_out/r02_cfg3.err; cut -c1-200 gpurun_out/r02_bench_cfg3_rpn_b256.json
timeout 1200 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu --no-context > gpurun_out/r02_bench_cfg5_s383_b128.json 2> gpurun_out/r02_cfg5.err
tail -2 gpurun_out/r02_cfg5.err; cut -c1-200 gpurun_out/r02_bench_cfg5_s383_b128.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_traffic.csv python bench.py --steps 6 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop > gpurun_out/r02_ncu_traffic.log 2>&1
tail -1 gpurun_out/r02_ncu_traffic.log | cut -c1-150; wc -l gpurun_out/r02_traffic.csv
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv3x3_patch_kernel<\(int\)64" -s 8 -c 2 -f -o gpurun_out/prof_patch64_r02 python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop > gpurun_out/r02_ncu_patch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel<\(int\)128, \(int\)2, \(int\)64, \(int\)1>" -s 60 -c 6 -f -o gpurun_out/prof_gemm_128_r02 python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop > gpurun_out/r02_ncu_gemm128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv_gemm_kernel<\(int\)256, \(int\)2, \(int\)64, \(int\)2>" -s 33 -c 4 -f -o gpurun_out/prof_gemm_pair_r02 python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu --no-context --no-verify --no-loop > gpurun_out/r02_ncu_pair.log 2>&1
ls -la gpurun_out/*_r02.ncu-rep
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "3x3_p1 or 1x1_64_256" > gpurun_out/r02_sanitizer_memcheck_ops.txt 2>&1
tail -4 gpurun_out/r02_sanitizer_memcheck_ops.txt
