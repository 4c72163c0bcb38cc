#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider -k "tile_variants or forced_wide or two_lane" 2>&1 | tail -4
