"""Copies the round-2 validation outputs from gpurun_out/ into profiles/ (tracked) and writes the derived summaries:
r02_traffic.json / r02_step_launches_traffic.csv (ncu DRAM bytes per launch of one step), r02_*_ncu_full.csv (condensed
ncu --set full captures), r02_other_configs.md."""
import json, os, shutil, subprocess, sys
G, P = "gpurun_out", "profiles"
def cp(src, dst=None):
    s = os.path.join(G, src)
    if os.path.exists(s) and os.path.getsize(s) > 0:
        shutil.copy(s, os.path.join(P, dst or src)); return True
    print("missing", src); return False
for f in ["r02_pytest_gpu.txt", "r02_bench_exact_n1.json", "r02_bench_fast_n1.json", "r02_bench_cfg3_rpn_b256.json",
          "r02_bench_cfg5_s383_b128.json", "r02_layers_exact.tsv", "r02_layers_fast.tsv", "r02_latency.log",
          "r02_sanitizer_memcheck_ops.txt", "r02_smoke.log"]:
    cp(f)
for n in (2, 4, 8):
    cp(f"r02_bench_exact_n{n}.json"); cp(f"r02_bench_cfg5_s383_n{n}.json")
if os.path.exists(os.path.join(G, "r02_traffic.csv")):
    subprocess.run([sys.executable, "tools/traffic_json.py", os.path.join(G, "r02_traffic.csv"),
                    os.path.join(P, "r02_traffic.json"), os.path.join(P, "r02_step_launches_traffic.csv")])
for rep, out in [("prof_patch64_r02", "r02_conv3x3_patch_ncu_full.csv"), ("prof_gemm_128_r02", "r02_conv_gemm_128_ncu_full.csv"),
                 ("prof_gemm_pair_r02", "r02_conv_gemm_pair_ncu_full.csv"), ("prof_xcorr_bulk_r02", "r02_xcorr_bulk_ncu_full.csv"),
                 ("prof_xcorr_nhwc_r02", "r02_xcorr_nhwc_ncu_full.csv")]:
    r = os.path.join(G, rep + ".ncu-rep")
    if os.path.exists(r):
        subprocess.run([sys.executable, "tools/ncu_summary.py", r, os.path.join(P, out)])
def load(f):
    try:
        return json.load(open(os.path.join(G, f)))
    except Exception:
        return None
lines = ["# Round-2 records for the other BASELINE configs and small batches (1xB200 unless noted; `tools/gpu_final_r02.sh`)", "",
         "| config | frames/s (device-resident, >= 2 s region) | ms/step | e2e frames/s | SM MHz | roofline frac (conv family) | parity_check max rel |",
         "|---|---|---|---|---|---|---|"]
for name, f in [("configs[1] B=64 sharp @255, exact", "r02_bench_exact_n1.json"), ("same, fast (single-pass fp16)", "r02_bench_fast_n1.json"),
                ("configs[2] SiamRPN-only B=256, exact", "r02_bench_cfg3_rpn_b256.json"),
                ("configs[4] search 383 B=128 sharp, exact", "r02_bench_cfg5_s383_b128.json")] + \
               [(f"configs[3] 64 streams/GPU x {n} GPUs", f"r02_bench_exact_n{n}.json") for n in (2, 4, 8)] + \
               [(f"configs[4] search 383, 128 streams/GPU x {n} GPUs", f"r02_bench_cfg5_s383_n{n}.json") for n in (2, 4, 8)]:
    d = load(f)
    if d is None:
        continue
    pc = d.get("parity_check") or {}
    rf = (d.get("roofline") or {}).get("frac")
    lines.append(f"| {name} | {d['value']:.0f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:.0f} | "
                 f"{(d.get('clocks') or {}).get('sm_mhz')} | {rf if rf is None else round(rf, 3)} | "
                 f"{pc.get('max_rel_all_ranks', pc.get('max_rel'))} |")
d = load("r02_bench_exact_n1.json")
if d:
    lines += ["", "Context on the same box (`context.cudnn`, the oracle port under PyTorch/cuDNN, frames/s): " + json.dumps(d.get("context", {}).get("cudnn")),
              "", "Whole tracker loop (`loop`): " + json.dumps(d.get("loop")),
              "", "CPU (`cpu_baseline`): " + json.dumps(d.get("cpu_baseline")),
              "", "Standalone xcorr: " + json.dumps(d.get("xcorr"))]
lat = os.path.join(G, "r02_latency.log")
if os.path.exists(lat):
    lines += ["", "Small batches (one whole frame per call, `tools/exp_latency.py`):", "", "```"] + open(lat).read().strip().splitlines() + ["```"]
open(os.path.join(P, "r02_other_configs.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
