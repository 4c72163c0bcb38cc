#!/bin/bash
mkdir -p gpurun_out
for prec in exact fast; do
timeout 600 python bench.py --rpn-only --batch 256 --steps 10 --warmup 3 --precision $prec > gpurun_out/cfg3_$prec.json 2> gpurun_out/cfg3_$prec.err; tail -2 gpurun_out/cfg3_$prec.err; cut -c1-330 gpurun_out/cfg3_$prec.json
timeout 600 python bench.py --search 383 --batch 128 --steps 6 --warmup 3 --no-cpu --precision $prec > gpurun_out/cfg5_$prec.json 2> gpurun_out/cfg5_$prec.err; tail -2 gpurun_out/cfg5_$prec.err; cut -c1-330 gpurun_out/cfg5_$prec.json
done
