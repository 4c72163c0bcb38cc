#!/usr/bin/env python
"""ncu long-format CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum of every kernel of a
`bench.py --steps 1 --warmup 3 --no-cpu` run, tools/gpu_traffic.sh) -> per-step DRAM traffic by kernel family.

One step = the launches from one lane-1 stem_tc launch to the next (ncu serialises launches in enqueue order: a step
enqueues lane 1's whole frame, then lane 0's; every lane's frame starts with stem_tc), i.e. two consecutive stem-to-stem
segments, taken from the 4th step after the template so that neither the template call nor the later profile /
end-to-end passes of bench.py are included."""
import csv
import json
import sys

FAMILIES = [("stem_tc", "stem_tc"), ("maxpool", "maxpool"), ("conv_gemm", "conv_gemm"), ("conv_patch", "conv3x3_patch"),
            ("xcorr", "xcorr"),
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
    # search-size stems only (grid 148 persistent; the 127 template stem has fewer tiles): every frame of a lane starts
    # with one
    stems = [i for i, k in enumerate(order) if "stem_tc" in launches[k]["kernel"]]
    assert len(stems) >= 12, "need a few steps of two lanes"
    first = stems[1]                      # stems[0] is the template pass
    seg = [launches[k] for k in order[stems[1 + 2 * 3]: stems[1 + 2 * 4]]]
    per = {}
    for name, pat in FAMILIES:
        ks = [l for l in seg if pat in l["kernel"]]
        per[name] = {"launches": len(ks),
                     "ncu_ms": round(sum(l["gpu__time_duration.sum"] for l in ks) / 1e6, 4),
                     "dram_read_MB": round(sum(l["dram__bytes_read.sum"] for l in ks) / 1e6, 1),
                     "dram_write_MB": round(sum(l["dram__bytes_write.sum"] for l in ks) / 1e6, 1)}
    g = [l for l in seg if "conv_gemm" in l["kernel"] or "conv3x3_patch" in l["kernel"]]
    res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control "
                     "none, one step of bench.py (B=64, search 255, exact; two lanes of 32 streams); conv_gemm_* totals "
                     "include the resident-patch 3x3 kernel",
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
