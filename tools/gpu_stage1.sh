#!/bin/bash
# First-contact run on the B200: each group in its own process with a timeout, logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to python -m pytest -q -rA -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/summary.txt
  grep -E "^\[parity\]|PASSED|FAILED|ERROR|passed|failed|Error|error" gpurun_out/$name.log | head -60 | tee -a gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run xcorr 300 tests/test_gpu_ops.py -k "xcorr"
run simt_conv 300 tests/test_gpu_ops.py -k "simt"
run tc_1x1 200 tests/test_gpu_ops.py -k "tensor_exact and 1x1_64_256"
run tc_im2col 200 tests/test_gpu_ops.py -k "tensor_exact and 3x3_p1 and not ds3"
run tc_all 600 tests/test_gpu_ops.py -k "tensor"
run engine_simt 900 tests/test_gpu_engine.py -k "simt"
run engine_tensor 900 tests/test_gpu_engine.py -k "not simt"
