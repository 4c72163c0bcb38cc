#!/bin/bash
# usage: gpurun --gpus N -- 'N=<N> bash tools/gpu_multi_r02.sh' : weak-scaling records for configs[1]/[3] (64 streams per GPU)
# and configs[4] (search 383, 128 streams per GPU) on N GPUs of one node
N=${N:-2}
mkdir -p gpurun_out
run() {  # outfile bench-args...
  out=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-context --no-loop "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  tail -2 gpurun_out/$out.err | cut -c1-200
  python - <<PY
import json
try:
    txt = open("gpurun_out/$out.json").read()
    line = [l for l in txt.splitlines() if l.startswith("{")][-1]
    open("gpurun_out/$out.json", "w").write(line + "\n")
    d = json.loads(line)
    print("$out", "N=", d["n_gpus"], round(d["value"]), "frames/s", round(d["ms_per_step"], 3), "ms/step; e2e", round(d["e2e"]["value"]),
          "; parity", d["parity_check"].get("max_rel_all_ranks", d["parity_check"]["max_rel"]), d["parity_check"].get("argmax_equal_all_ranks"))
except Exception as e:
    print("$out failed", e)
PY
}
run r02_bench_exact_n$N
run r02_bench_cfg5_s383_n$N --config 5
