#!/bin/bash
# A/B: exact-mode 128x256 tiles on the 1x1 layers (SMB200_EXACT_N256) — per-layer tables + parity check.
mkdir -p gpurun_out
for v in 0 1 2 3; do
  SMB200_EXACT_N256=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --dump-layers gpurun_out/layers_n256_$v.tsv \
     > gpurun_out/bench_n256_$v.json 2> gpurun_out/bench_n256_$v.err
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/bench_n256_$v.json"))
    print("n256=$v", r["value"], r["ms_per_step"], r.get("e2e", {}).get("value"))
except Exception as e:
    print("n256=$v failed", e)
PY
done
SMB200_EXACT_N256=1 timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
