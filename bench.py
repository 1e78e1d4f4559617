#!/usr/bin/env python
"""Benchmark of the local-ICP point-to-point matching hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A step = one pass of the whole hot path (landmark partition + bin sort, all partition ICPs, fill, blend, final
map, both directions) over one batch: F maize-sized frame pairs (BASELINE configs[1]: ~50k points per frame,
12 long ribbon segments each).  F = 64 by default so that the float64 inputs (~150 MB) exceed the 126 MB L2 and
successive steps cannot reuse them from cache.

`value`    device-resident throughput: inputs already in HBM, CUDA events on the launching stream.
`e2e`      the same metric through the C-ABI call with HOST buffers (pinned): H2D, kernels and D2H inside, pipelined
           over 8 sub-batches by liblocalicp.
`roofline` the dominant kernel against the FP32 pipe (NN kernels are FP32-bound, 8 flop per
           evaluated source/candidate pair; `traffic` = its DRAM bytes per launch from the committed ncu capture);
           `roofline_hbm` the whole step against measured HBM with the algorithmic bytes of SURVEY.md section 8d.
Under torchrun every rank runs its own F frame pairs (weak scaling, no data-path collective).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "local_icp point correspondences/sec"
UNIT = "correspondences/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE config number (1..5); 2 = maize (default)")
    ap.add_argument("--frame-pairs", type=int, default=64, help="frame pairs per step and per GPU")
    ap.add_argument("--cpu-sample-frame-pairs", type=int, default=0, help="0 = auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--map-mode", type=int, default=0, help="final map: 0 = landmark-bin tiles (default), 1 = hashed grid")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm, one process per host core over segment pairs
# ------------------------------------------------------------------------------------------------
def _cpu_one(seg):
    from oracle import local_icp_oracle as O
    x1, x2, l1, l2 = seg
    # as the reference runs: sklearn KDTree + numpy SVD + per-point Python blend loop
    T21, T12 = O.segment_pair_maps(x1, x2, l1, l2, nn="kdtree", loop=True)
    return len(T21) + len(T12)


def cpu_throughput(batch, n_proc, repeats=1):
    """correspondences/s of the CPU port on `batch` using n_proc worker processes"""
    import multiprocessing as mp
    segs = batch.segments()
    os.environ["OMP_NUM_THREADS"] = "1"
    ctx = mp.get_context("fork")
    with ctx.Pool(n_proc) as pool:
        pool.map(_cpu_one, segs[:n_proc])                  # warm the workers (imports)
        best = None
        for _ in range(repeats):
            t = time.perf_counter()
            n = sum(pool.map(_cpu_one, segs, chunksize=1))
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
    return n / best, best, n


def cpu_sample_frame_pairs(args, cores):
    """bounded CPU sample: about 10-30 s of CPU work.  One core does roughly 0.4 maize/tomato frame pairs per second
    (SURVEY.md section 6), i.e. ~2.5 core-seconds per frame pair; config 5 is ~110 core-seconds per frame pair."""
    per_fp = {1: 2.5, 2: 2.7, 3: 8.8, 4: 2.5 * 15, 5: 110.0}[args.config]
    return int(max(1, min(args.frame_pairs, round(20.0 / per_fp))))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi samples of SM clock / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """one process per GPU: run (and first-touch the pinned staging buffers) on the CPU cores next to that GPU, so
    that N concurrent host-to-device streams do not all cross the same socket link"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def make_workload(args, rank):
    syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
    # frame index offset per rank: every rank registers different frame pairs
    b = syn.make_config(args.config, frame_pairs=args.frame_pairs, frame_offset=rank * args.frame_pairs)
    b.name = f"C{args.config}x{args.frame_pairs}"
    return b


def workload_name(args):
    names = {1: "tomato", 2: "maize", 3: "arabidopsis", 4: "tomato series", 5: "1M-point"}
    return (f"BASELINE configs[{args.config - 1}] ({names[args.config]}-sized frame pair), batch of {args.frame_pairs} "
            f"frame pairs per step per GPU, both directions (T21+T12)")


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    a2 = argparse.Namespace(**vars(args))
    a2.frame_pairs = args.cpu_sample_frame_pairs or cpu_sample_frame_pairs(args, cores)
    batch = make_workload(a2, 0)
    n_proc = min(cores, batch.n_segments)
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
    segs = batch.segments()
    times = []
    with mp.get_context("fork").Pool(n_proc) as pool:
        for i in range(args.warmup + args.steps):
            t = time.perf_counter()
            n = sum(pool.map(_cpu_one, segs, chunksize=1))
            dt = time.perf_counter() - t
            if i >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = n / (ms / 1e3)
    sample = f"{a2.frame_pairs} frame pair(s) = {batch.n_segments} segment pairs, {n} correspondences per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": sample},
            "frame_pairs_per_s": a2.frame_pairs / (ms / 1e3),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": n_proc, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_b200(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    native = importlib.import_module("4d_plant_analysis_b200._native")
    if world > 1:
        bind_to_gpu_numa_node(local)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = native.Context(local)
    ctx.set_map_mode(args.map_mode)
    batch = make_workload(args, rank)
    ctx.plan(batch.off1, batch.off2, batch.lm_off)
    ncorr = batch.n_correspondences
    # ---- device-resident inputs / outputs (torch owns the memory; liblocalicp gets raw pointers) ----
    d_in = [torch.from_numpy(a).to(dev) for a in (batch.xyz1, batch.xyz2, batch.lm1, batch.lm2)]
    d_T21 = torch.empty(int(batch.off2[-1]), dtype=torch.int64, device=dev)
    d_T12 = torch.empty(int(batch.off1[-1]), dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream(dev)          # a real (non-null) stream: handle 0 would select the library's own
    torch.cuda.set_stream(stream)

    def step_device():
        ctx.run_device(d_in[0].data_ptr(), d_in[1].data_ptr(), d_in[2].data_ptr(), d_in[3].data_ptr(),
                       d_T21.data_ptr(), d_T12.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.counters()["launches"] * args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    # ---- per-kernel times: identical steps, CUDA events recorded between the kernels on the same stream ----
    ctx.set_profiling(True)
    stage = {}
    for _ in range(max(3, min(args.steps, 10))):
        step_device()
        torch.cuda.synchronize(dev)
        for k, v in ctx.stage_ms().items():
            stage.setdefault(k, []).append(v)
    counters = ctx.counters()
    pair_counts = ctx.pair_counts()
    ctx.set_profiling(False)
    stage = {k: float(np.mean(v)) for k, v in stage.items()}
    tr = ctx.transforms(filled=False)
    bin_off = ctx.bin_offsets()
    # ---- end to end through the C-ABI host entry point, pinned host buffers ----
    pins = [torch.from_numpy(a).pin_memory() for a in (batch.xyz1, batch.xyz2, batch.lm1, batch.lm2)]
    h_T21 = torch.empty(int(batch.off2[-1]), dtype=torch.int64).pin_memory()
    h_T12 = torch.empty(int(batch.off1[-1]), dtype=torch.int64).pin_memory()

    def step_host():
        ctx.run_host_ptr(pins[0].data_ptr(), pins[1].data_ptr(), pins[2].data_ptr(), pins[3].data_ptr(),
                         h_T21.data_ptr(), h_T12.data_ptr())

    for _ in range(3):
        step_host()
    barrier()
    th0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - th0)
    assert torch.equal(h_T21.to(dev), d_T21) and torch.equal(h_T12.to(dev), d_T12), "host and device paths disagree"
    h2d = sum(int(p.numel()) * 8 for p in pins)
    d2h = int(h_T21.numel() + h_T12.numel()) * 8
    # ---- max over ranks, whole-job totals ----
    vals = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(ncorr), float(batch.n_frame_pairs)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms = float(vals[0]), float(vals[1])
    ncorr_all, nfp_all = float(tot[0]), float(tot[1])
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = ncorr_all / (ms_step / 1e3)
    # ---- roofline of the dominant kernel (this rank's numbers) ----
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    sumK = ctx.sumK
    sizes = np.diff(bin_off).astype(np.float64)
    ns = np.concatenate([sizes[sumK:], sizes[:sumK]])          # rows in "21 then 12" order: src sizes
    nt = np.concatenate([sizes[:sumK], sizes[sumK:]])
    iters = np.concatenate([tr["iters21"], tr["iters12"]]).astype(np.float64)
    mx = np.maximum(ns, nt)
    w = iters * ns * nt                                        # evaluated (source, candidate) pairs per partition
    split = 98304 if 2 * sumK >= 12000 else 32768               # LIP_WARP_WORK_LARGE / LIP_WARP_WORK of the (unsplit) profiled run
    brute = {"icp_group_kernel": float(w[ns * nt <= split].sum()),          # n_src * n_tgt * passes: what a brute-force scan evaluates
             "icp_big_kernel": float(w[ns * nt > split].sum())}
    # algorithmic pair counts: brute force inside a partition for the ICP kernels (padding excluded), the candidates the
    # final map actually evaluates (box-pruned bins, windowed along the sort axis)
    pairs = dict(brute, final_map_kernel=float(pair_counts["final_map_kernel"]))
    big = ns * nt > split
    it_stats = {"group": {"n": int((~big & (iters > 0)).sum()), "mean": float(iters[~big & (iters > 0)].mean()), "p99": float(np.percentile(iters[~big & (iters > 0)], 99)), "max": float(iters[~big].max())},
                "big": {"n": int((big & (iters > 0)).sum()), "mean": float(iters[big & (iters > 0)].mean()) if (big & (iters > 0)).any() else 0.0,
                        "p99": float(np.percentile(iters[big & (iters > 0)], 99)) if (big & (iters > 0)).any() else 0.0, "max": float(iters[big].max()) if big.any() else 0.0}}
    dom = max(pairs, key=lambda k: stage[k])
    fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12       # TFLOP/s, FP32 FMA pipe at the max SM clock
    achieved = 8.0 * pairs[dom] / (stage[dom] / 1e3) / 1e12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if os.path.exists(tpath) and args.config == 2 and args.frame_pairs == 64 and args.map_mode == 0:   # captured on the default workload only
        tj = json.load(open(tpath))
        kmap = {"final_map_kernel": ("blend_tile_kernel",)}
        names = kmap.get(dom, (dom,))
        if all(n in tj["kernels"] for n in names):
            traffic = sum(tj["kernels"][n]["dram_bytes_per_launch"] for n in names)
            traffic_src = "profiles/" + tj["source"]
    roof = {"kernel": dom, "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved / fp32_peak, "traffic": traffic, "traffic_unit": "dram bytes per launch (ncu --set full)",
            "traffic_source": traffic_src, "kernel_ms": stage[dom],
            "pairs_per_launch": pairs[dom], "flop_per_pair": 8, "brute_force_pairs": brute,
            "peak_source": f"148 SM x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz ({peak_src})",
            "per_kernel": {k: {"ms": stage[k], "tflops": 8.0 * pairs[k] / (stage[k] / 1e3) / 1e12 if stage[k] > 0 else 0.0,
                               "pairs": pairs[k]} for k in pairs}}
    n_pts = float(batch.off1[-1] + batch.off2[-1])
    mean_it = float(iters[iters > 0].mean())
    alg_bytes = (48.0 * mean_it + 56.0) * n_pts / 2.0          # (48 I + 56) N per frame pair, N = points per frame
    step_ms_this = sum(stage.values())
    roof_hbm = {"bound": "hbm", "scope": "whole step (all kernels)", "achieved": alg_bytes / (ms_step / 1e3) / 1e9 if world == 1
                else alg_bytes / (step_ms_this / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "algorithmic_bytes_per_step": alg_bytes, "mean_nn_passes": mean_it, "peak_source": peak_src}
    roof_hbm["frac"] = roof_hbm["achieved"] / hbm_peak
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "segment_pairs_per_gpu": batch.n_segments,
                       "points_per_gpu": int(n_pts), "correspondences_per_step": int(ncorr_all),
                       "l2": "inputs per step (%.0f MB float64 + workspace) exceed the 126 MB L2; no flush" % (h2d / 1e6),
                       "arithmetic": "fp32 NN search with float64 re-evaluation of near ties; float64 sums, SVD, transforms"},
            "frame_pairs_per_s": nfp_all / (ms_step / 1e3),
            "e2e": {"value": ncorr_all / (e2e_ms / args.steps / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "frame_pairs_per_s": nfp_all / (e2e_ms / args.steps / 1e3)},
            "gpu_launches": int(launches), "kernels_per_step": counters["launches"],
            "stage_ms": stage, "counters": counters, "nn_passes": it_stats, "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm}
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        a2 = argparse.Namespace(**vars(args))
        a2.frame_pairs = args.cpu_sample_frame_pairs or cpu_sample_frame_pairs(args, cores)
        sample = make_workload(a2, 0)
        n_proc = min(cores, sample.n_segments)
        v, dt, n = cpu_throughput(sample, n_proc)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": n_proc, "kind": "port",
                                "sample": f"{a2.frame_pairs} frame pair(s) of the same workload = {sample.n_segments} segment pairs, "
                                          f"{n} correspondences, {dt:.1f} s wall on {n_proc} processes "
                                          "(oracle port: sklearn KDTree + numpy SVD + per-point blend loop, as the reference runs)"}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
