#!/bin/bash
# r02 tile A/B after the convergent-issue fix
mkdir -p gpurun_out
run() {  # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --no-verify --min-seconds 1 \
      --dump-layers gpurun_out/r2e_layers_$name.tsv > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2e_$name.json"))
    print("$name", round(d["value"]), round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"], round(d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e)
PY
}
run default SMB200_LANES=2
run wide16 SMB200_WIDE_KB=16
run wide6 SMB200_WIDE_KB=6
run wide16_nopair SMB200_WIDE_KB=16 SMB200_CTA_PAIR=0
run wide6_nopair SMB200_WIDE_KB=6 SMB200_CTA_PAIR=0
run pair2 SMB200_CTA_PAIR=2
run n256_0 SMB200_EXACT_N256=0
run lanes1 SMB200_LANES=1
run lanes3 SMB200_LANES=3
( timeout 600 python -m pytest tests/test_batch_tracker.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/r2e_tracker.log 2>&1
