#!/bin/bash
# deconv with an in-block K split: refine parity (engine tests that reach the refine module) + single-stream latency
mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider -k "sharp_b1_matches_oracle or reference_golden or batched_streams" 2>&1 | tail -4 ) > gpurun_out/r2p_engine.log 2>&1
tail -2 gpurun_out/r2p_engine.log
SMB200_LAT_ONLY=exact1 timeout 60 python tools/exp_latency.py > gpurun_out/r2p_latency.log 2>&1
cat gpurun_out/r2p_latency.log
