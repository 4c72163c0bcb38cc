#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log 2>&1
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --dump-layers gpurun_out/layers_exact.tsv > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
tail -3 gpurun_out/bench_exact.err; cut -c1-300 gpurun_out/bench_exact.json
timeout 900 python bench.py --steps 20 --warmup 3 --precision fast --no-cpu --dump-layers gpurun_out/layers_fast.tsv > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
tail -3 gpurun_out/bench_fast.err; cut -c1-300 gpurun_out/bench_fast.json
