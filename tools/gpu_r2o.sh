#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "tile_variants" 2>&1 | tail -5 ) > gpurun_out/r2o_variants.log 2>&1
tail -3 gpurun_out/r2o_variants.log
