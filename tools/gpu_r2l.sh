#!/bin/bash
# N-concatenated exact MMA (A_hi x [B_hi;B_lo]) + deconv load pipelining: parity, then A/B bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py tests/test_gpu_shapes.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/r2l_tests.log 2>&1
tail -3 gpurun_out/r2l_tests.log
for v in 0 1; do
  SMB200_NO_NCAT=$v timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --no-loop --dump-layers gpurun_out/r2l_layers_ncat$v.tsv > gpurun_out/r2l_bench_noncat$v.json 2> gpurun_out/r2l_bench_noncat$v.err
  tail -2 gpurun_out/r2l_bench_noncat$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_noncat$v.json')); print('NO_NCAT=$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['kernels_ms_per_step']['conv_gemm']['ms'], d['kernels_ms_per_step']['refine_misc']['ms'], d['parity_check']['max_rel'])"
done
SMB200_NO_NCAT=0 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-context --no-loop --no-verify > gpurun_out/r2l_bench_noncat0_b.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_noncat0_b.json')); print('NO_NCAT=0 again', round(d['value']), d['clocks']['sm_mhz'])"
