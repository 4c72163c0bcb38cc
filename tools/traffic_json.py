#!/usr/bin/env python
"""ncu long-format CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum of every kernel of a
`bench.py --steps 1 --warmup 3 --no-cpu` run, tools/gpu_traffic.sh) -> per-step DRAM traffic by kernel family.

One step = the launches between two consecutive select_kernel launches (refine of step k, track of step k+1: the same
set of kernels as one step), taken between the 2nd and 3rd select so that neither the template call nor the later
profile / end-to-end passes of bench.py are included."""
import csv
import json
import sys

FAMILIES = [("stem_tc", "stem_tc"), ("maxpool", "maxpool"), ("conv_gemm", "conv_gemm"), ("xcorr", "xcorr"),
            ("select", "select_kernel"), ("crop", "crop"), ("small_conv", "small_conv"), ("gather", "gather_corr"),
            ("deconv", "deconv")]


def main(src, out_json, out_csv=None):
    rows = []
    with open(src) as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    launches = {}
    order = []
    for r in csv.DictReader(lines[start:]):
        k = int(r["ID"])
        if k not in launches:
            launches[k] = {"kernel": r["Kernel Name"], "grid": r["Grid Size"]}
            order.append(k)
        launches[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    sel = [i for i, k in enumerate(order) if "select_kernel" in launches[k]["kernel"]]
    assert len(sel) >= 3, "need at least three select_kernel launches"
    seg = [launches[k] for k in order[sel[1] + 1: sel[2] + 1]]
    per = {}
    for name, pat in FAMILIES:
        ks = [l for l in seg if pat in l["kernel"]]
        per[name] = {"launches": len(ks),
                     "ncu_ms": round(sum(l["gpu__time_duration.sum"] for l in ks) / 1e6, 4),
                     "dram_read_MB": round(sum(l["dram__bytes_read.sum"] for l in ks) / 1e6, 1),
                     "dram_write_MB": round(sum(l["dram__bytes_write.sum"] for l in ks) / 1e6, 1)}
    g = [l for l in seg if "conv_gemm" in l["kernel"]]
    res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control "
                     "none, one step of bench.py (B=64, search 255, exact; two lanes of 32 streams)",
           "per_step": per,
           "conv_gemm_traffic_bytes_per_step": sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in g),
           "conv_gemm_launches_per_step": len(g)}
    json.dump(res, open(out_json, "w"), indent=1)
    if out_csv:
        with open(out_csv, "w") as f:
            f.write("idx,kernel,grid,duration_ns,dram_read_bytes,dram_write_bytes\n")
            for i, l in enumerate(seg):
                f.write(f'{i},{l["kernel"].replace(",", ";")[:100]},{l["grid"].replace(",", ";")},'
                        f'{l["gpu__time_duration.sum"]:.0f},{l["dram__bytes_read.sum"]:.0f},{l["dram__bytes_write.sum"]:.0f}\n')
    print(json.dumps({k: v for k, v in res.items() if k != "per_step"}), len(seg), "launches in the step")


if __name__ == "__main__":
    main(*sys.argv[1:])
