#!/bin/bash
# CTA-pair (cta_group::2) GEMM bring-up: op parity, engine parity, per-layer A/B.
mkdir -p gpurun_out
export SMB200_CTA_PAIR=2
if timeout 300 python -m pytest tests/test_gpu_ops.py -k conv -q -x -p no:cacheprovider > gpurun_out/pair_ops.log 2>&1; then
  tail -3 gpurun_out/pair_ops.log
  timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_tracker_loop.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pair_engine.log
  cat gpurun_out/pair_engine.log
  for v in 0 1 2; do
    SMB200_CTA_PAIR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --dump-layers gpurun_out/layers_pair_$v.tsv \
       > gpurun_out/bench_pair_$v.json 2> gpurun_out/bench_pair_$v.err
    python -c "
import json
try:
    r = json.load(open('gpurun_out/bench_pair_$v.json')); print('pair=$v exact', r['value'], r['ms_per_step'])
except Exception as e: print('pair=$v failed', e)"
  done
  for v in 0 1; do
    SMB200_CTA_PAIR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --precision fast --dump-layers gpurun_out/layers_pairfast_$v.tsv \
       > gpurun_out/bench_pairfast_$v.json 2> gpurun_out/bench_pairfast_$v.err
    python -c "
import json
try:
    r = json.load(open('gpurun_out/bench_pairfast_$v.json')); print('pair=$v fast', r['value'], r['ms_per_step'])
except Exception as e: print('pair=$v fast failed', e)"
  done
else
  tail -30 gpurun_out/pair_ops.log
  for k in 1x1_64_256 3x3_d2_p2 3x3_p0_kernel 3x3_v2 1x1_mask3969; do
    echo "== $k"; timeout 120 python -m pytest tests/test_gpu_ops.py -k "exact and $k" -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|error|max_rel|rel" | head -5
  done
  echo "== fast 3x3_d2_p2"; timeout 120 python -m pytest tests/test_gpu_ops.py -k "fast and 3x3_d2_p2" -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|Error|error|rel" | head -5
fi
