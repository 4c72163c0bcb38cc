#!/bin/bash
# Full GPU test suite + smoke + bench + ncu launch list.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log 2>&1
cat gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
tail -3 gpurun_out/bench_exact.err; cat gpurun_out/bench_exact.json | cut -c1-3000
timeout 900 python bench.py --steps 10 --warmup 3 --precision fast --no-cpu > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
tail -3 gpurun_out/bench_fast.err; cat gpurun_out/bench_fast.json | cut -c1-3000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
