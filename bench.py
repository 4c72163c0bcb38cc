#!/usr/bin/env python
"""Benchmark of the SiamMask per-frame inference hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = one pass of the hot path over one batch of synthetic search regions:
`track_mask` (backbone -> depthwise xcorr -> cls/loc/mask heads) + `track_refine` for B=64 paired tracker
streams per GPU (BASELINE.json configs[1]: "batch=64 synthetic search regions, 1xB200, full track() path with
mask refine"); templates are cached per slot (configs[3]).  N>1: one process per GPU (torchrun), streams are
sharded, the packed weights are broadcast ONCE over NCCL at init, no per-frame collective ("weak" scaling).

Prints ONE JSON line (rank 0).  `value` = whole-job frames/s with inputs resident in HBM; `e2e` = the same
metric through the C-ABI host-buffer call (H2D of every frame + D2H of cls/loc/mask logits inside the timed
region); `roofline` = the tensor-core conv family (dominant kernel) timed per launch with CUDA events;
`cpu_baseline` = the oracle port of the reference timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json

import numpy as np
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "search-region frames/sec (127/255 SiamMask-sharp)"
GFLOP_PER_FRAME = {255: 33.915, 383: 77.938}          # BASELINE.md §2 (algorithmic, conv_kernel cached)
XCORR_BYTES = {255: 1526784, 383: 3820544}            # per branch per frame, fp32 algorithmic


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The process is
    started (and its first sample awaited) before the warm-up, so its start-up cost never lands inside the timed
    region; a reader thread time-stamps every sample and `stop()` keeps those inside [t0, t1]."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, uuid):
        import threading
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.first = threading.Event()

        def reader():
            for line in self.proc.stdout:
                self.lines.append((time.perf_counter(), line))
                self.first.set()
        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()
        self.first.wait(5.0)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        self.thread.join(5.0)
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.lines:
            if not (t0 <= ts <= t1 + 0.03):
                continue
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference algorithm on the host CPU: the oracle port (oracle/siammask_oracle.py), all host threads."""
    if rank != 0:
        return
    import torch
    from oracle.siammask_oracle import Oracle
    from siammask_b200.checkpoint import synthetic_state_dict
    sd = synthetic_state_dict(0)
    bs = args.ref_batch
    g = torch.Generator().manual_seed(1)
    z = torch.rand(bs, 3, 127, 127, generator=g) * 255
    xs = [torch.rand(bs, 3, args.search, args.search, generator=g) * 255 for _ in range(2)]
    pos = torch.randint(0, 25, (bs, 2), generator=g).numpy()
    o = Oracle(sd)
    o.template(z)

    def step(i):
        o.track_mask(xs[i % 2])
        o.track_refine(pos)
    cores = pick_threads(lambda: step(0), torch)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    fps = bs * args.steps / dt
    sample = (f"{args.steps} steps x {bs} paired frames, track_mask+track_refine, torch CPU fp32, {cores} threads "
              f"(fastest of 8/16/32/64/all {os.cpu_count()})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, bs, 1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, batch_per_gpu, world):
    return {"workload": f"SiamMask-sharp config_davis, template 127 / search {args.search}, response "
                        f"{(args.search - 127) // 8 + 9}x{(args.search - 127) // 8 + 9}: track_mask (incl. 256->3969 "
                        f"mask head) + on-device score/box selection + track_refine at the selected position, {batch_per_gpu} paired streams per GPU, templates cached per slot",
            "global_batch": batch_per_gpu * world, "batch_per_gpu": batch_per_gpu, "search": args.search,
            "parallelism": f"streams sharded over {world} GPU(s), one NCCL weight broadcast at init, "
                           "no per-frame collective; inside a GPU the batch runs as two concurrent lanes of "
                           "batch_per_gpu/2 streams (batches >= 16)",
            "l2": "inputs rotate over 4 device buffers (4 x 50 MB) and every step streams > 5 GB of activations "
                  "(>> 126 MB L2)"}


def pick_threads(fn, torch):
    """torch's CPU convs stop scaling (and collapse) well before 100+ threads at batch 1: try a few thread
    counts for ~1 s each and keep the fastest, so the CPU baseline is the best the host can do."""
    ncpu = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for n in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
        if dt > 3.0:
            break
    torch.set_num_threads(best)
    return best


def cpu_baseline(args, seconds=10.0):
    import torch
    from oracle.siammask_oracle import Oracle
    from siammask_b200.checkpoint import synthetic_state_dict
    sd = synthetic_state_dict(0)
    g = torch.Generator().manual_seed(1)
    z = torch.rand(1, 3, 127, 127, generator=g) * 255
    x = torch.rand(1, 3, args.search, args.search, generator=g) * 255
    o = Oracle(sd)
    o.template(z)

    def one():
        o.track_mask(x); o.track_refine((12, 12))
    cores = pick_threads(one, torch)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        o.track_mask(x); o.track_refine((12, 12))
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{n} frames, B=1 track_mask+track_refine((12,12)), oracle port (torch CPU fp32, "
                      f"{cores} threads = fastest of 8/16/32/64/all), {dt:.1f} s"}


def run_gpu(args, rank, local_rank, world):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import siammask_b200 as smb
    from siammask_b200 import _lib
    from siammask_b200.parallel import broadcast_weights, max_over_ranks, shard_streams

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = args.search
    # weak scaling: args.batch streams per GPU; this rank owns a contiguous block of the global stream ids
    B = len(shard_streams(args.batch * world, world, rank))
    R = (S - 127) // 8 + 9
    m = smb.Custom(anchors=smb.DEFAULT_ANCHORS, search_size=S, max_batch=B, num_slots=B, precision=args.precision,
                   mask=not args.rpn_only)
    if rank == 0:
        m.load_state_dict(smb.synthetic_state_dict(0, mask=not args.rpn_only, refine=not args.rpn_only))
    m.eval().to(dev)
    if world > 1:                       # the one collective of the whole job: weights, once, at init
        broadcast_weights(m.weight_blob(), src=0)
        torch.cuda.synchronize()
        if rank != 0:
            m.adopt_weights()
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    z = torch.rand(B, 3, 127, 127, device=dev, generator=gen) * 255
    xs = [torch.rand(B, 3, S, S, device=dev, generator=gen) * 255 for _ in range(4)]
    pos = torch.randint(0, R, (B, 2), device=dev, generator=gen, dtype=torch.int32)
    # what siamese_init prepares per stream (tools/test.py:142-161): anchors, cosine window, target size in the crop
    from siammask_b200 import tracker
    anchors_dev = torch.from_numpy(tracker.generate_anchor(smb.DEFAULT_ANCHORS, R)).to(dev)
    window_dev = torch.from_numpy(np.tile(np.outer(np.hanning(R), np.hanning(R)).flatten(), 5).astype(np.float32)).to(dev)
    tsz_dev = torch.rand(B, 2, device=dev, generator=gen) * 60 + 30
    m.template(z)

    def step_rpn(i, mask_head=True):
        cls, loc = m.track(xs[i % 4])
        return m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4)

    def step(i, mask_head=True):
        if args.rpn_only:
            return step_rpn(i)
        # the full per-frame path of siamese_track (tools/test.py:201-261) without leaving the device:
        # track_mask -> score/box post-processing + argmax -> track_refine at the selected position
        cls, loc, mask = m.track_mask(xs[i % 4], mask_head=mask_head)
        best, sel_pos, rec = m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4)
        ref = m.track_refine(sel_pos)
        return (cls, loc, mask, rec), ref

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), device=dev)

    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        step(i)
    l0 = m.launch_count
    t_start = time.perf_counter()
    ms = timed(step, args.steps)
    t_end = time.perf_counter()
    launches = m.launch_count - l0
    clocks = sampler.stop(t_start, t_end) if sampler else None
    fps = world * B * args.steps / (ms * 1e-3)
    ms_skip = timed(lambda i: step(i, mask_head=False), args.steps)
    fps_skip = world * B * args.steps / (ms_skip * 1e-3)

    if args.rpn_only:
        if rank == 0:
            gfl = {255: 30.811, 383: 71.139}.get(S, 0.0)
            print(json.dumps({"metric": METRIC.replace("SiamMask-sharp", "SiamRPN-only"), "value": fps, "unit": "frames/s",
                              "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                              "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "precision_mode": args.precision, "data": "synthetic",
                              "config": {"workload": f"SiamRPN-only track + on-device selection, search {S}, {B} streams/GPU"},
                              "algorithmic_tflops": fps * gfl / 1e3, "gpu_launches": launches, "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of x, track + refine, D2H of results
    lib = _lib.load()
    A = 5
    xh = [torch.empty(B, 3, S, S).pin_memory() for _ in range(2)]
    for t in xh:
        t.copy_(xs[0].cpu())
    clsh = [torch.empty(B, 2 * A, R, R).pin_memory() for _ in range(2)]
    loch = [torch.empty(B, 4 * A, R, R).pin_memory() for _ in range(2)]
    maskh = [torch.empty(B, 127 * 127).pin_memory() for _ in range(2)]
    posh = pos.cpu().contiguous().pin_memory()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def submit(i):
        tk = C.c_int32()
        _lib.check(lib.sm_track_host_async(m.handle, 0, B, xh[i % 2].data_ptr(), clsh[i % 2].data_ptr(),
                                           loch[i % 2].data_ptr(), posh.data_ptr(), maskh[i % 2].data_ptr(), stream,
                                           C.byref(tk)))
        return tk.value

    def host_loop(n):
        # a user with a stream of frames keeps one step in flight: submit frame k+1, then collect frame k
        prev = None
        for i in range(n):
            tk = submit(i)
            if prev is not None:
                _lib.check(lib.sm_track_host_wait(m.handle, prev))
            prev = tk
        _lib.check(lib.sm_track_host_wait(m.handle, prev))
    host_loop(3)
    barrier()
    t0 = time.perf_counter()
    host_loop(args.steps)      # every step's results are in host memory when this returns
    dt = max_over_ranks(time.perf_counter() - t0, device=dev)
    e2e_fps = world * B * args.steps / dt
    h2d = xh[0].numel() * 4 + posh.numel() * 4
    d2h = (clsh[0].numel() + loch[0].numel() + maskh[0].numel()) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-launch CUDA-event timing of every kernel (same workload, separate pass)
    peaks = load_peaks()
    m.profile(True)
    nprof = 3
    for i in range(nprof):
        step(i)
    rows = m.profile_dump()
    m.profile(False)
    cats = {}
    for name, cat, t, fl, by in rows:
        c = cats.setdefault(cat, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        c["ms"] += t / nprof; c["flops"] += fl / nprof; c["bytes"] += by / nprof; c["launches"] += 1 / nprof
    gemm = cats.get("conv_gemm", {"ms": 1e-9, "flops": 0.0, "bytes": 0.0, "launches": 0})
    tot_ms = sum(c["ms"] for c in cats.values())
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12
    layers = {}
    for name, cat, t, fl, by in rows:
        if cat == "conv_gemm":
            L = layers.setdefault(name, [0.0, 0.0])
            L[0] += t / nprof; L[1] += fl / nprof
    top = sorted(layers.items(), key=lambda kv: -kv[1][0])[:6]
    if args.dump_layers:
        with open(args.dump_layers, "w") as f:
            f.write("name\tcat\tms\tgflop\tMB\tTFLOPs\tGBps\n")
            for name, cat, t, fl, by in rows[:len(rows) // nprof]:
                f.write(f"{name}\t{cat}\t{t:.4f}\t{fl / 1e9:.2f}\t{by / 1e6:.1f}\t{fl / (t * 1e-3) / 1e12:.1f}\t"
                        f"{by / (t * 1e-3) / 1e9:.0f}\n")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.precision == "exact" and B == 64 and S == 255:
        tj = json.load(open(tpath))           # ncu dram__bytes_read+write of the family's launches in one step
        traffic = {"bytes_per_step": tj["conv_gemm_traffic_bytes_per_step"],
                   "launches_per_step": tj["conv_gemm_launches_per_step"], "source": "profiles/r01_traffic.json (ncu)"}
    roofline = {
        "bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv family, all layers of one step)",
        "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
        "peak_source": peaks["src"] + ", sustained cuBLAS bf16", "traffic": traffic,
        "algorithmic_bytes_per_step": gemm["bytes"],
        # the parity mode issues 3 fp16 MMAs per algorithmic MAC: tensor-pipe work actually executed vs the same peak
        "mma_issued_frac": (3.0 if args.precision == "exact" else 1.0) * achieved / peaks["tflops"],
        "launches_per_step": gemm["launches"], "ms_per_step": gemm["ms"], "share_of_step": gemm["ms"] / tot_ms,
        "algorithmic_gflop_per_step": gemm["flops"] / 1e9,
        "note": "algorithmic FLOPs (2*M*N*K per conv, no padding, no x3 for the split-fp16 passes) / summed "
                "CUDA-event durations of the launches",
        "top_layers": [{"layer": k, "ms": v[0], "tflops": v[1] / (v[0] * 1e-3) / 1e12} for k, v in top],
    }
    by_cat = {k: {"ms": round(v["ms"], 4), "launches": round(v["launches"], 1),
                  "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else 0.0,
                  "gbps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0}
              for k, v in sorted(cats.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- standalone depthwise xcorr operator (the "xcorr GB/s" half of the metric): 3 branches x 64 streams
    planes_b = 3 * B
    xc = torch.randn(planes_b, 256, R + 4, R + 4, device=dev)
    kc = torch.randn(planes_b, 256, 5, 5, device=dev)
    for _ in range(3):
        smb.conv2d_dw_group(xc, kc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 10
    e0.record()
    for _ in range(reps):
        out = smb.conv2d_dw_group(xc, kc)
    e1.record()
    torch.cuda.synchronize()
    xms = e0.elapsed_time(e1) / reps
    xbytes = planes_b * XCORR_BYTES.get(S, (256 * ((R + 4) ** 2 + 25 + R * R)) * 4)
    xgbs = xbytes / (xms * 1e-3) / 1e9
    del xc, kc, out

    result = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f16x3 (hi+lo split fp16 operands on tcgen05, f32 accumulate; f32 CUDA-core stem/xcorr/refine)"
                 if args.precision == "exact" else "f16 (single-pass tcgen05, f32 accumulate)",
        "data": "synthetic", "config": workload_config(args, B, world),
        "precision_mode": args.precision,
        "algorithmic_tflops": fps * GFLOP_PER_FRAME.get(S, 0.0) / 1e3,
        "value_skip_dead_mask_head": fps_skip,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "sm_track_host_async / sm_track_host_wait (C ABI, pinned host buffers, one step in flight; "
                       "refine positions supplied by the host; mask head skipped as under --refine)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernels_ms_per_step": by_cat,
        "xcorr": {"op": "sm_xcorr_depthwise fp32 NCHW", "planes": planes_b * 256, "ms": xms, "GBps": xgbs,
                  "peak": peaks["hbm_gbs"], "frac": xgbs / peaks["hbm_gbs"], "peak_source": peaks["src"]},
        "device_bytes": m.device_bytes,
    }
    if world == 1 and not args.no_cpu:
        result["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="paired tracker streams per GPU")
    ap.add_argument("--search", type=int, default=255)
    ap.add_argument("--precision", default="exact", choices=["exact", "fast"])
    ap.add_argument("--rpn-only", action="store_true",
                    help="SiamRPN-only engine (experiments/siamrpn_resnet): step = track -> cls/loc (BASELINE configs[2])")
    ap.add_argument("--ref-batch", type=int, default=2, help="frames per step of the CPU reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dump-layers", default=None, help="write the per-launch CUDA-event table of one step here")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
