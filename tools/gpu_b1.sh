#!/bin/bash
mkdir -p gpurun_out
for b in 1 4 8; do
for v in 1 3; do
SMB200_EXACT_N256=$v timeout 300 python bench.py --batch $b --steps 200 --warmup 10 --no-cpu > gpurun_out/b${b}_n256_$v.json 2> gpurun_out/b${b}_n256_$v.err
python -c "
import json
r=json.load(open('gpurun_out/b${b}_n256_$v.json')); print('B=$b n256=$v', round(r['value']), round(r['ms_per_step'],4), 'e2e', round(r['e2e']['value']))"
done
done
