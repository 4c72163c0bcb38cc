#!/bin/bash
# after the cross-term reorder: whole GPU suite, smoke, headline bench
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/r2m_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2m_pytest_gpu.txt
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/r2m_smoke.log 2>&1
tail -2 gpurun_out/r2m_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --dump-layers gpurun_out/r2m_layers.tsv > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -2 gpurun_out/r2m_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench.json')); print(round(d['value']), round(d['value_skip_dead_mask_head']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['roofline']['frac'], d['kernels_ms_per_step'], d['parity_check']['max_rel'])"
timeout 600 python tools/exp_latency.py > gpurun_out/r2m_latency.log 2>&1
cat gpurun_out/r2m_latency.log
