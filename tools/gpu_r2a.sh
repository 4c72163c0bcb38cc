#!/bin/bash
# r02 first GPU pass: new shape tests first, then the whole GPU suite, smoke, a default bench run.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/r2a_shapes.log 2>&1
tail -5 gpurun_out/r2a_shapes.log
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_shapes.py 2>&1 | tail -30 ) > gpurun_out/r2a_pytest.log 2>&1
tail -3 gpurun_out/r2a_pytest.log
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/r2a_smoke.log 2>&1
tail -2 gpurun_out/r2a_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 --dump-layers gpurun_out/r2a_layers.tsv > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_bench.err; cut -c1-400 gpurun_out/r2a_bench.json
