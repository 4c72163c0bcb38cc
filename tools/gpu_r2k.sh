#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider -k "sharp_b1_matches_oracle or reference_golden or batched_streams or search_383" 2>&1 | tail -8 ) > gpurun_out/r2k_engine.log 2>&1
tail -3 gpurun_out/r2k_engine.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --dump-layers gpurun_out/r2k_layers.tsv > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -2 gpurun_out/r2k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['kernels_ms_per_step']['stem'], d['parity_check']['max_rel'])"
timeout 600 python bench.py --batch 1 --steps 20 --warmup 3 --min-seconds 0.2 --no-cpu --no-context --no-loop --dump-layers gpurun_out/r2k_layers_b1.tsv > gpurun_out/r2k_bench_b1.json 2> gpurun_out/r2k_bench_b1.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_b1.json')); print('B=1', round(d['value']), round(d['ms_per_step'],3), d['kernels_ms_per_step'])"
