#!/bin/bash
# N-lane engine: parity + A/B over the lane count
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_tracker_loop.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/lanes_tests.log
cat gpurun_out/lanes_tests.log
SMB200_LANES=4 timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider -k "two_lane or host or graph" 2>&1 | tail -3
for v in 2 3 4 1; do
  for prec in exact fast; do
    SMB200_LANES=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --precision $prec \
       > gpurun_out/bench_lanes${v}_$prec.json 2> gpurun_out/bench_lanes${v}_$prec.err
    python -c "
import json
try:
    r = json.load(open('gpurun_out/bench_lanes${v}_$prec.json')); print('lanes=$v $prec', round(r['value']), r['ms_per_step'], 'e2e', round(r['e2e']['value']), 'skip', round(r.get('value_skip_dead_mask_head')))
except Exception as e: print('lanes=$v $prec failed', e)"
  done
done
tail -5 gpurun_out/bench_lanes2_exact.err
