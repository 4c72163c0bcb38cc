#!/bin/bash
# quick check of the xcorr load pipelining + loop leg + engine tests
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_engine.py tests/test_gpu_shapes.py tests/test_gpu_range.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r2g_engine.log 2>&1
tail -4 gpurun_out/r2g_engine.log
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --dump-layers gpurun_out/r2g_layers.tsv > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -3 gpurun_out/r2g_bench.err; cut -c1-300 gpurun_out/r2g_bench.json
grep -E "corr_" gpurun_out/r2g_layers.tsv | cut -f1,3,7
