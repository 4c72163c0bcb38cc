#!/bin/bash
# full GPU suite on the final build + the other BASELINE configs
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/smoke.log 2>&1
cat gpurun_out/smoke.log
for prec in exact fast; do
timeout 600 python bench.py --rpn-only --batch 256 --steps 10 --warmup 3 --precision $prec > gpurun_out/cfg3_$prec.json 2> gpurun_out/cfg3_$prec.err; tail -2 gpurun_out/cfg3_$prec.err; cut -c1-330 gpurun_out/cfg3_$prec.json
timeout 600 python bench.py --search 383 --batch 128 --steps 6 --warmup 3 --no-cpu --precision $prec > gpurun_out/cfg5_$prec.json 2> gpurun_out/cfg5_$prec.err; tail -2 gpurun_out/cfg5_$prec.err; cut -c1-330 gpurun_out/cfg5_$prec.json
done
timeout 300 python bench.py --batch 1 --steps 200 --warmup 10 --no-cpu > gpurun_out/b1_exact.json 2> gpurun_out/b1_exact.err; cut -c1-250 gpurun_out/b1_exact.json
