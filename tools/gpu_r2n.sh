#!/bin/bash
# tile thresholds again after the concatenated cross term made the 128-wide tiles cheaper
mkdir -p gpurun_out
for v in "SMB200_WIDE_KB=24" "SMB200_WIDE_KB=40" "SMB200_WIDE_KB=80" "SMB200_CTA_PAIR=2" "SMB200_WIDE_KB=24"; do
  tag=$(echo $v | tr '=' '_')
  env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --no-loop --no-verify --dump-layers gpurun_out/r2n_layers_$tag.tsv > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/r2n_$tag.json')); print('$v', round(d['value']), round(d['value_skip_dead_mask_head']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['kernels_ms_per_step']['conv_gemm']['ms'])"
done
