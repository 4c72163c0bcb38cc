#!/bin/bash
# two-lane engine: parity + A/B against the single-lane schedule
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_tracker_loop.py tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/lanes_tests.log
cat gpurun_out/lanes_tests.log
for v in 2 1; do
  for prec in exact fast; do
    SMB200_LANES=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --precision $prec \
       > gpurun_out/bench_lanes${v}_$prec.json 2> gpurun_out/bench_lanes${v}_$prec.err
    python -c "
import json
try:
    r = json.load(open('gpurun_out/bench_lanes${v}_$prec.json')); print('lanes=$v $prec', round(r['value']), r['ms_per_step'], 'e2e', round(r['e2e']['value']), 'skip', r.get('value_skip_dead_mask_head'))
except Exception as e: print('lanes=$v $prec failed', e)"
  done
done
tail -5 gpurun_out/bench_lanes2_exact.err
