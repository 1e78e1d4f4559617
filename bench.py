#!/usr/bin/env python
"""Benchmark of the local-ICP point-to-point matching hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A step = one pass of the whole hot path (landmark partition + bin sort, all partition ICPs, fill, blend, final map, both
directions) over one batch.  Headline workload (every N): a batch of growth series of maize-sized frames (BASELINE configs[1]
frame pairs, batched as configs[3] asks: all consecutive frame pairs of a series) — 4 plants x 17 frames = 64 frame pairs,
~7.9 M correspondences, per GPU (weak scaling: every rank registers its own plants).  Every frame is stored and uploaded
ONCE although it is day2 of one frame pair and day1 of the next (run_registration_pipeline.py:64,100-104).

`value`     device-resident throughput: inputs already in HBM, CUDA events on the launching stream.
`e2e`       the same metric through the C-ABI call with pinned HOST buffers (H2D, kernels, D2H inside, pipelined over
            sub-batches by liblocalicp); `e2e.variants` times the other host paths (pageable numpy, new shape + plan, the
            round-1 style batch that uploads shared frames twice and returns int64 maps).
`roofline`  the dominant kernel against the FP32 pipe (NN kernels are FP32-bound, 8 flop per evaluated
            source/candidate pair); `roofline_hbm` the whole step against measured HBM with the algorithmic bytes of section 8d.
`parity`    the oracle (CPU restatement of the reference) on random segment pairs of the batch that was just timed.
`configs`   (N = 1) one entry per BASELINE config: C1x64, C2x64 (independent frame pairs), C3x16, C4 (16-frame series), C5.
`sharded`   the multi-GPU data path of the design on the configs BASELINE names for it: C4 (15 frame pairs dealt 2,2,2,2,2,2,2,1)
            and C5 (one 1M-point frame pair, segment pairs dealt by cost), STRONG scaling, through ShardedLocalICP with the
            NCCL gather of maps + score counts inside the timed region and gathered == single-GPU asserted in the run.
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
NAMES = {1: "tomato", 2: "maize", 3: "arabidopsis", 4: "tomato series", 5: "1M-point"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="frame size of the headline series: BASELINE config 1, 2 (default, maize) or 3")
    ap.add_argument("--plants", type=int, default=4, help="growth series per GPU")
    ap.add_argument("--frames", type=int, default=17, help="frames per series (frame pairs per series = frames - 1)")
    ap.add_argument("--independent", action="store_true", help="headline = independent frame pairs (round-1 workload) instead of series")
    ap.add_argument("--frame-pairs", type=int, default=64, help="frame pairs per step and per GPU with --independent")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config table (N = 1)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded C4 / C5 runs")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--int32-maps", action="store_true", help="headline with int32 maps instead of uint16")
    ap.add_argument("--parity-pairs", type=int, default=16)
    ap.add_argument("--map-mode", type=int, default=0, help="final map: 0 = landmark-bin tiles (default), 1 = hashed grid, 2 = tiles with tensor-core scores")
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


def _cpu_parity(seg):
    from oracle import local_icp_oracle as O
    x1, x2, l1, l2 = seg
    d21, d12 = {}, {}
    T21 = O.segment_map(x1, x2, l1, l2, nn="kdtree", details=d21)
    T12 = O.segment_map(x2, x1, l2, l1, nn="kdtree", details=d12)
    keep = ("R", "t", "valid", "iters")
    return T21, T12, {k: d21[k] for k in keep}, {k: d12[k] for k in keep}


def cpu_throughput(segs, n_proc):
    """correspondences/s of the CPU port on `segs` using n_proc worker processes (1: in this process, serial — how the
    reference actually runs)"""
    os.environ["OMP_NUM_THREADS"] = "1"
    if n_proc == 1:
        _cpu_one(segs[0])
        t = time.perf_counter()
        n = sum(_cpu_one(s) for s in segs)
        return n / (time.perf_counter() - t), time.perf_counter() - t, n
    import multiprocessing as mp
    with mp.get_context("fork").Pool(n_proc) as pool:
        pool.map(_cpu_one, segs[:n_proc])                  # warm the workers (imports)
        t = time.perf_counter()
        n = sum(pool.map(_cpu_one, segs, chunksize=1))
        dt = time.perf_counter() - t
    return n / dt, dt, n


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def syn():
    return importlib.import_module("4d_plant_analysis_b200.synthetic")


def headline_workload(args, rank):
    if args.independent:
        b = syn().make_config(args.config, frame_pairs=args.frame_pairs, frame_offset=rank * args.frame_pairs)
        name = (f"BASELINE configs[{args.config - 1}] ({NAMES[args.config]}-sized frame pair), batch of {args.frame_pairs} independent "
                f"frame pairs per step per GPU, both directions (T21+T12)")
        return b, name
    s = syn().make_series(args.config, n_frames=args.frames, plants=args.plants, seed_offset=rank * args.plants)
    name = (f"BASELINE configs[{args.config - 1}] ({NAMES[args.config]}-sized) frames batched as growth series (configs[3] style): "
            f"{args.plants} plants x {args.frames} frames = {s.n_frame_pairs} consecutive frame pairs per step per GPU, both "
            f"directions (T21+T12); every frame stored and uploaded once")
    return s, name


# ------------------------------------------------------------------------------------------------
# one workload on one GPU: device-resident steps, per-kernel times, end-to-end steps
# ------------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, torch, native, local, data, *, map_mode=0, map_dtype=np.int32):
        self.torch, self.data = torch, data
        self.dev = torch.device("cuda", local)
        self.series = hasattr(data, "frames")
        self.ctx = native.Context(local)
        self.ctx.set_map_mode(map_mode)
        self.ctx.set_map_dtype(map_dtype)
        self.map_dtype = map_dtype
        self.tdt = {np.uint16: torch.int16, np.int32: torch.int32, np.int64: torch.int64}[map_dtype]   # (uint16 maps live in int16 storage)
        self.ctx.plan(data.off1, data.off2, data.lm_off)
        self.host = [data.frames, data.lm1, data.lm2] if self.series else [data.xyz1, data.xyz2, data.lm1, data.lm2]
        self.d_in = [torch.from_numpy(a).to(self.dev) for a in self.host]
        self.n21, self.n12 = int(data.off2[-1]), int(data.off1[-1])
        self.d_T21 = torch.empty(self.n21, dtype=self.tdt, device=self.dev)
        self.d_T12 = torch.empty(self.n12, dtype=self.tdt, device=self.dev)
        self.stream = torch.cuda.Stream(self.dev)      # a real (non-null) stream: handle 0 would select the library's own
        self.pins = None

    def step_device(self):
        p = [t.data_ptr() for t in self.d_in]
        if self.series:
            self.ctx.run_series_device(p[0], len(self.data.frames), self.data.row1, self.data.row2, p[1], p[2],
                                       self.d_T21.data_ptr(), self.d_T12.data_ptr(), self.stream.cuda_stream)
        else:
            self.ctx.run_device(p[0], p[1], p[2], p[3], self.d_T21.data_ptr(), self.d_T12.data_ptr(), self.stream.cuda_stream)

    def pin(self):
        torch = self.torch
        if self.pins is None:
            self.pins = [torch.from_numpy(a).pin_memory() for a in self.host]
            self.h_T21 = torch.empty(self.n21, dtype=self.tdt).pin_memory()
            self.h_T12 = torch.empty(self.n12, dtype=self.tdt).pin_memory()

    def step_host(self):
        p = [t.data_ptr() for t in self.pins]
        if self.series:
            self.ctx.run_series_ptr(p[0], len(self.data.frames), self.data.row1, self.data.row2, p[1], p[2],
                                    self.h_T21.data_ptr(), self.h_T12.data_ptr())
        else:
            self.ctx.run_host_ptr(p[0], p[1], p[2], p[3], self.h_T21.data_ptr(), self.h_T12.data_ptr())

    @property
    def h2d_bytes(self):
        return sum(int(a.size) * 8 for a in self.host)

    @property
    def d2h_bytes(self):
        return (self.n21 + self.n12) * np.dtype(self.map_dtype).itemsize

    def time_device(self, steps, warmup, barrier):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            self.step_device()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            self.step_device()
        e1.record(self.stream)
        barrier()
        return e0.elapsed_time(e1) / steps

    def time_host(self, steps, warmup, barrier):
        self.pin()
        for _ in range(max(warmup, 3)):
            self.step_host()
        barrier()
        t = time.perf_counter()
        for _ in range(steps):
            self.step_host()
        self.torch.cuda.synchronize(self.dev)
        ms = 1e3 * (time.perf_counter() - t) / steps
        ok = self.torch.equal(self.h_T21.to(self.dev), self.d_T21) and self.torch.equal(self.h_T12.to(self.dev), self.d_T12)
        assert ok, "host and device paths disagree"
        return ms

    def profile(self, reps):
        """identical steps with CUDA events between the kernels (kernels serialised, batch in one piece)"""
        self.ctx.set_profiling(True)
        stage = {}
        for _ in range(reps):
            self.step_device()
            self.torch.cuda.synchronize(self.dev)
            for k, v in self.ctx.stage_ms().items():
                stage.setdefault(k, []).append(v)
        counters, pair_counts = self.ctx.counters(), self.ctx.pair_counts()
        self.ctx.set_profiling(False)
        stage = {k: float(np.mean(v)) for k, v in stage.items()}
        tr = self.ctx.transforms(filled=False)
        bin_off = self.ctx.bin_offsets()
        return stage, counters, pair_counts, tr, bin_off

    def close(self):
        self.ctx.close()


def roofline_of(run, stage, pair_counts, tr, bin_off, sm_max_mhz, peak_src):
    """dominant NN kernel against the FP32 pipe: 8 flop per evaluated (source, candidate) pair"""
    sumK = run.ctx.sumK
    sizes = np.diff(bin_off).astype(np.float64)
    ns = np.concatenate([sizes[sumK:], sizes[:sumK]])          # rows in "21 then 12" order: src sizes
    nt = np.concatenate([sizes[:sumK], sizes[sumK:]])
    iters = np.concatenate([tr["iters21"], tr["iters12"]]).astype(np.float64)
    w = iters * ns * nt                                        # evaluated (source, candidate) pairs per partition
    split = 98304 if 2 * sumK >= 12000 else 32768              # LIP_WARP_WORK_LARGE / LIP_WARP_WORK of the (unsplit) profiled run
    big = ns * nt > split
    brute = {"icp_group_kernel": float(w[~big].sum()), "icp_big_kernel": float(w[big].sum())}   # what a brute-force scan evaluates
    pairs = dict(brute, final_map_kernel=float(pair_counts["final_map_kernel"]))
    ran = iters > 0
    it_stats = {}
    for name, sel in (("group", ~big & ran), ("big", big & ran)):
        it_stats[name] = {"n": int(sel.sum()), "mean": float(iters[sel].mean()) if sel.any() else 0.0,
                          "p99": float(np.percentile(iters[sel], 99)) if sel.any() else 0.0, "max": float(iters[sel].max()) if sel.any() else 0.0}
    dom = max(pairs, key=lambda k: stage[k])
    fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12        # TFLOP/s, FP32 FMA pipe at the max SM clock
    per_kernel = {k: {"ms": stage[k], "tflops": 8.0 * pairs[k] / (stage[k] / 1e3) / 1e12 if stage[k] > 0 else 0.0, "pairs": pairs[k],
                      "frac": (8.0 * pairs[k] / (stage[k] / 1e3) / 1e12 / fp32_peak) if stage[k] > 0 else 0.0} for k in pairs}
    roof = {"kernel": dom, "bound": "fp32", "achieved": per_kernel[dom]["tflops"], "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": per_kernel[dom]["frac"], "traffic": None, "kernel_ms": stage[dom], "pairs_per_launch": pairs[dom], "flop_per_pair": 8,
            "peak_source": f"148 SM x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz ({peak_src})", "per_kernel": per_kernel}
    return roof, it_stats, float(iters[ran].mean()) if ran.any() else 0.0


def parity_check(args, data, run, tr_filled, tr_raw):
    """the oracle on random segment pairs of the batch that was timed; GPU maps from the last end-to-end step"""
    rng = np.random.default_rng(12345)
    S = data.n_segments
    pick = np.sort(rng.choice(S, size=min(args.parity_pairs, S), replace=False))
    batch = data.as_batch() if hasattr(data, "frames") else data
    segs = [tuple(np.array(a) for a in batch.segment(int(s))) for s in pick]
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
    t = time.perf_counter()
    with mp.get_context("fork").Pool(min(host_cores(), len(segs))) as pool:
        ref = pool.map(_cpu_parity, segs, chunksize=1)
    dt = time.perf_counter() - t
    g21, g12 = run.h_T21.numpy(), run.h_T12.numpy()
    if run.map_dtype == np.uint16:
        g21, g12 = g21.view(np.uint16), g12.view(np.uint16)
    mism = tot = parts = same_it = 0
    max_dR = max_dt = 0.0
    for s, (T21, T12, d21, d12) in zip(pick, ref):
        a, b = int(data.off2[s]), int(data.off2[s + 1])
        c, d = int(data.off1[s]), int(data.off1[s + 1])
        mism += int((g21[a:b] != T21).sum() + (g12[c:d] != T12).sum())
        tot += (b - a) + (d - c)
        e, f = int(data.lm_off[s]), int(data.lm_off[s + 1])
        for det, R, tt, it, st in ((d21, tr_filled["R21"], tr_filled["t21"], tr_raw["iters21"], tr_raw["status21"]),
                                   (d12, tr_filled["R12"], tr_filled["t12"], tr_raw["iters12"], tr_raw["status12"])):
            ok = st[e:f] == 1                                  # ran and every solve had full rank (R is unique there)
            parts += int(ok.sum())
            same_it += int((it[e:f][ok] == det["iters"][ok]).sum())
            if ok.any():
                max_dR = max(max_dR, float(np.abs(R[e:f][ok] - det["R"][ok]).max()))
                max_dt = max(max_dt, float(np.abs(tt[e:f][ok] - det["t"][ok]).max()))
    return {"checked": int(len(pick)), "segment_pairs_in_batch": int(S), "correspondences_checked": int(tot), "map_mismatch": int(mism),
            "partitions_checked": int(parts), "iter_identical_frac": same_it / max(parts, 1), "max_dR": max_dR, "max_dt": max_dt,
            "oracle": "oracle.local_icp_oracle (sklearn KDTree + numpy SVD), %.1f s" % dt,
            "tolerance": "maps identical up to exact float64 ties; R, t within 1e-5 (BASELINE north_star)"}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """the reference's CPU algorithm on the headline workload: every host core on a bounded sample of the same generator"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    # a bounded sample of the headline generator: frame pairs of single plants, truncated to a multiple of the core count (at
    # least two segment pairs per core) so that no core idles while the others finish
    per_fp = len(syn().make_series(args.config, n_frames=2, plants=1, seed_offset=100).as_batch().segments())
    n_plants = max(1, -(-2 * cores // per_fp))
    segs = syn().make_series(args.config, n_frames=2, plants=n_plants, seed_offset=100).as_batch().segments()
    if len(segs) >= cores:
        segs = segs[:len(segs) // cores * cores]
    n_proc = min(cores, len(segs))
    import multiprocessing as mp
    os.environ["OMP_NUM_THREADS"] = "1"
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
    v1, dt1, n1 = cpu_throughput(segs[:12], 1)
    sample = f"{len(segs)} segment pairs of the same generator ({n} correspondences) per step on {n_proc} processes"
    full, full_name = headline_workload(args, 0)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": full_name, "sample": sample, "host_cores": cores, "cpu": cpu_model(),
                       "note": "throughput-normalised: the CPU arm times a bounded sample of the workload the GPU arm runs in full"},
            "frame_pairs_per_s": val / (full.n_correspondences / full.n_frame_pairs),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": n_proc, "kind": "port", "sample": sample,
                             "ref_1core": {"value": v1, "cores": 1, "sample": f"12 segment pairs, {n1} correspondences, {dt1:.1f} s, serial in one process "
                                                                              "(how the reference actually runs)"}},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def sharded_run(torch, dist, D, native, local, rank, world, data, steps, barrier, label):
    """strong scaling of ONE batch over the ranks: device-resident and end-to-end step times (gather inside), checked against
    the whole batch on one GPU"""
    dev = torch.device("cuda", local)
    sh = D.ShardedLocalICP(data, device=local)
    for _ in range(3):
        out = sh.step()
    sh.stream.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sh.stream)
    for _ in range(steps):
        out = sh.step()
    e1.record(sh.stream)
    sh.stream.synchronize()
    barrier()
    dev_ms = e0.elapsed_time(e1) / steps
    for _ in range(2):
        out = sh.step_from_host()
    sh.stream.synchronize()
    barrier()
    t = time.perf_counter()
    for _ in range(steps):
        out = sh.step_from_host()
        if rank == 0:                                            # the whole-batch result lands in rank 0's host memory
            h = [o.cpu() for o in out]
    sh.stream.synchronize()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t) / steps
    # gathered == single GPU (rank 0 runs the whole batch alone, outside the timed regions)
    ok = True
    if rank == 0:
        ref = Runner(torch, native, local, data, map_dtype=np.int64)
        ref.step_device()
        ref.stream.synchronize()
        ok = bool(torch.equal(out[0].to(torch.int64), ref.d_T21) and torch.equal(out[1].to(torch.int64), ref.d_T12))
        ref.close()
    vals = torch.tensor([dev_ms, e2e_ms, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    n = data.n_correspondences
    counts = out[2].cpu().numpy()
    res = {"workload": label, "scaling": "strong", "segment_pairs": int(data.n_segments), "frame_pairs": int(data.n_frame_pairs),
           "correspondences": int(n), "ms_per_step": float(vals[0]), "value": n / (float(vals[0]) / 1e3),
           "e2e_ms_per_step": float(vals[1]), "e2e_value": n / (float(vals[1]) / 1e3),
           "gathered_equals_single_gpu": bool(float(vals[2]) == 0.0),
           "shard_segment_pairs": [int(len(s)) for s in sh.plan.seg_ids],
           "gather": "one all_gather_into_tensor of %d int32 per rank (maps of both directions + score counts), count vector exchanged at plan time" % sh.plan.pad,
           "boundary_frame_rows_over_nvlink": int(sh.boundary_rows),
           "mean_scores": [float(x) for x in D.scores_from_counts(counts).mean(axis=0)]}
    sh.ctx.close()
    return res


def run_b200(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    native = importlib.import_module("4d_plant_analysis_b200._native")
    D = importlib.import_module("4d_plant_analysis_b200.distributed")
    if world > 1:
        bind_to_gpu_numa_node(local)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    data, wname = headline_workload(args, rank)
    # the maps are indices into one segment's mesh: uint16 holds them whenever every mesh has fewer than 65,536 vertices
    small = max(int(np.diff(data.off1).max()), int(np.diff(data.off2).max())) < 65536
    run = Runner(torch, native, local, data, map_mode=args.map_mode, map_dtype=np.uint16 if small and not args.int32_maps else np.int32)
    ncorr = data.n_correspondences
    # ---- device-resident steps, clocks sampled during the timed region ----
    for _ in range(max(args.warmup, 3)):
        run.step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t0 = time.perf_counter()
    ms_step = run.time_device(args.steps, 0, barrier)
    t1 = time.perf_counter()
    launches = run.ctx.counters()["launches"] * args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    stage, counters, pair_counts, tr_raw, bin_off = run.profile(max(3, min(args.steps, 10)))
    # ---- end to end through the C-ABI host entry point, pinned host buffers ----
    e2e_ms = run.time_host(args.steps, 3, barrier)
    tr_filled = run.ctx.transforms(filled=True)
    tr_raw_e2e = run.ctx.transforms(filled=False)
    # ---- max over ranks, whole-job totals ----
    vals = torch.tensor([ms_step, e2e_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(ncorr), float(data.n_frame_pairs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step_all, e2e_ms_all = float(vals[0]), float(vals[1])
    ncorr_all, nfp_all = float(tot[0]), float(tot[1])
    # ---- the multi-GPU data path of the design: C4 by frame pair, C5 by cost, strong scaling, gather inside ----
    sharded = None
    if not args.no_sharded:
        sharded = {}
        c4 = syn().make_series(4, n_frames=16)
        sharded["c4"] = sharded_run(torch, dist, D, native, local, rank, world, c4, max(5, args.steps), barrier,
                                    "BASELINE configs[3]: 16-frame tomato growth series, 15 consecutive frame pairs, sharded by frame pair")
        c5 = syn().make_config(5)
        sharded["c5"] = sharded_run(torch, dist, D, native, local, rank, world, c5, max(5, args.steps), barrier,
                                    "BASELINE configs[4]: one 1M-point frame pair, 2,000 segment pairs, sharded by cost (longest processing time first)")
    if rank != 0:
        run.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    roof, it_stats, mean_it = roofline_of(run, stage, pair_counts, tr_raw, bin_off, sm_max_mhz, peak_src)
    tpath = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if os.path.exists(tpath) and not args.independent and args.config == 2 and args.plants == 4 and args.frames == 17:
        tj = json.load(open(tpath))                              # captured on the default workload only
        names = {"final_map_kernel": "blend_tile_kernel"}
        k = names.get(roof["kernel"], roof["kernel"])
        if k in tj["kernels"]:
            roof["traffic"] = tj["kernels"][k]["dram_bytes_per_launch"]
            roof["traffic_unit"] = "dram bytes per launch (ncu --set full)"
            roof["traffic_source"] = "profiles/" + tj["source"]
    n_pts = float(data.off1[-1] + data.off2[-1])
    alg_bytes = (48.0 * mean_it + 56.0) * n_pts / 2.0            # (48 I + 56) N per frame pair, N = points per frame
    roof_hbm = {"bound": "hbm", "scope": "whole step (all kernels)", "achieved": alg_bytes / (ms_step / 1e3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "algorithmic_bytes_per_step": alg_bytes, "mean_nn_passes": mean_it, "peak_source": peak_src}
    roof_hbm["frac"] = roof_hbm["achieved"] / hbm_peak
    value = ncorr_all / (ms_step_all / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step_all, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wname, "segment_pairs_per_gpu": data.n_segments, "points_per_gpu": int(n_pts),
                       "correspondences_per_step": int(ncorr_all), "host_cores": host_cores(),
                       "maps": ("uint16" if run.map_dtype == np.uint16 else "int32") + " (lip_set_map_dtype: entries are indices into one segment's mesh, every mesh here has "
                               "fewer than 65,536 vertices; the drop-in functions return int64 like the reference)",
                       "l2": "inputs per step (%.0f MB float64 + workspace) exceed the 126 MB L2 only together with the ~1 GB workspace the kernels stream; no flush" % (run.h2d_bytes / 1e6),
                       "arithmetic": "fp32 NN search with float64 re-evaluation of near ties; float64 sums, SVD, transforms"},
            "frame_pairs_per_s": nfp_all / (ms_step_all / 1e3),
            "e2e": {"value": ncorr_all / (e2e_ms_all / 1e3), "unit": UNIT, "h2d_bytes_per_step": run.h2d_bytes,
                    "d2h_bytes_per_step": run.d2h_bytes, "ms_per_step": e2e_ms_all, "frame_pairs_per_s": nfp_all / (e2e_ms_all / 1e3),
                    "path": ("lip_series_run_host" if run.series else "lip_batch_run_host") + ", pinned host buffers, " + ("uint16" if run.map_dtype == np.uint16 else "int32") + " maps"},
            "gpu_launches": int(launches), "kernels_per_step": counters["launches"],
            "stage_ms": stage, "counters": counters, "nn_passes": it_stats, "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm}
    # ---- parity inside the run: the oracle on random segment pairs of the batch just timed ----
    if not args.no_parity:
        assert np.array_equal(tr_raw_e2e["iters21"], tr_raw["iters21"]), "pipelined and one-piece runs disagree on iteration counts"
        line["parity"] = parity_check(args, data, run, tr_filled, tr_raw_e2e)
    # ---- the other host paths of the same workload (this rank only) ----
    if world == 1:
        line["e2e"]["variants"] = e2e_variants(torch, native, local, args, data, run, ncorr)
        line["body_e2e"] = body_e2e(native, local, args, data, ncorr)
    run.close()
    # ---- one entry per BASELINE config ----
    if world == 1 and not args.no_configs:
        line["configs"] = configs_table(torch, native, local, args, sm_max_mhz, peak_src, barrier)
    if sharded is not None:
        line["sharded"] = sharded
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_variants(torch, native, local, args, data, run, ncorr):
    """host-to-host times of the same workload through the other entry points a caller might use"""
    out = {}
    steps = max(3, min(args.steps, 5))
    batch = data.as_batch() if run.series else data
    # (a) pageable numpy arrays, as a caller of the reference holds them; same map type as the headline
    ctx = run.ctx
    fn = (lambda: ctx.run_series(data.frames, data.row1, data.row2, data.lm1, data.lm2)) if run.series else \
         (lambda: ctx.run_host(data.xyz1, data.xyz2, data.lm1, data.lm2))
    for _ in range(2):
        fn()
    t = time.perf_counter()
    for _ in range(steps):
        fn()
    ms = 1e3 * (time.perf_counter() - t) / steps
    out["pageable_numpy"] = {"ms_per_step": ms, "value": ncorr / (ms / 1e3), "path": "Context.run_series / run_host on ordinary numpy arrays (staged through pinned buffers by the library's copy threads)"}
    # (b) the reference-shaped batched call: validation on, plan of a NEW shape every call, pageable arrays, int64 maps
    B = importlib.import_module("4d_plant_analysis_b200.batch")
    c2 = native.Context(local)
    S = batch.n_segments
    subs = [batch.take(np.arange(0, S - k)) for k in (0, 1, 2)]
    B.local_icp_batch(subs[0], ctx=c2)
    times = []
    for sb in subs[1:] + subs[:1]:
        t = time.perf_counter()
        B.local_icp_batch(sb, ctx=c2)
        times.append(1e3 * (time.perf_counter() - t))
    ms = float(np.mean(times))
    out["pageable_new_shape_validated"] = {"ms_per_step": ms, "value": ncorr / (ms / 1e3),
                                           "path": "local_icp_batch(): lip_batch_plan of a shape not seen before + NaN/inf validation + lip_batch_run_host on pageable arrays, int64 maps, shared frames uploaded twice"}
    # (b2) the headline path with int32 maps
    if run.map_dtype == np.uint16:
        r32 = Runner(torch, native, local, data, map_dtype=np.int32)
        r32.step_device()
        ms = r32.time_host(steps, 2, lambda: torch.cuda.synchronize())
        out["pinned_int32_maps"] = {"ms_per_step": ms, "value": ncorr / (ms / 1e3), "d2h_bytes_per_step": r32.d2h_bytes, "path": "the headline path with int32 maps"}
        r32.close()
    # (c) round-1 style: gathered batch (every shared frame uploaded twice), pinned, int64 maps
    r1 = Runner(torch, native, local, batch, map_dtype=np.int64)
    r1.step_device()
    ms = r1.time_host(steps, 2, lambda: torch.cuda.synchronize())
    out["pinned_batch_int64"] = {"ms_per_step": ms, "value": ncorr / (ms / 1e3), "h2d_bytes_per_step": r1.h2d_bytes, "d2h_bytes_per_step": r1.d2h_bytes,
                                 "path": "lip_batch_run_host on the gathered batch, pinned buffers, int64 maps (the round-1 bench path)"}
    r1.close()
    c2.close()
    return out


def body_e2e(native, local, args, data, ncorr):
    """the whole per-segment body of the reference's driver (p2p_matching_seg_icp.py:173-189: both maps, double_direction_refinement,
    evaluate_p2p_matching) for the bench batch in ONE call from ordinary numpy arrays (lip_batch_match_host)"""
    B = importlib.import_module("4d_plant_analysis_b200.batch")
    batch = data.as_batch() if hasattr(data, "frames") else data
    ctx = native.Context(local)
    B.match_segments_batch(batch, ctx=ctx)
    times = []
    for _ in range(3):
        t = time.perf_counter()
        r = B.match_segments_batch(batch, ctx=ctx)
        times.append(1e3 * (time.perf_counter() - t))
    ctx.close()
    ms = float(np.median(times))
    return {"ms_per_step": ms, "value": ncorr / (ms / 1e3), "unit": UNIT, "frame_pairs_per_s": batch.n_frame_pairs / (ms / 1e3),
            "mean_scores": [float(x) for x in np.asarray(r.score)],
            "path": "match_segments_batch(): T21 + T12, double_direction_refinement, evaluate_p2p_matching of every segment pair, pageable numpy in, numpy out"}


def configs_table(torch, native, local, args, sm_max_mhz, peak_src, barrier):
    """C1x64, C2x64 (independent frame pairs), C3x16, C4 (16-frame series), C5: device-resident and end-to-end throughput and the
    roofline fraction of the dominant kernel"""
    rows = []
    steps = max(3, min(args.steps, 5))
    for cfg, fp, series in ((1, 64, False), (2, 64, False), (3, 16, False), (4, 15, True), (5, 1, False)):
        data = syn().make_series(4, n_frames=16) if series else syn().make_config(cfg, frame_pairs=fp)
        r = Runner(torch, native, local, data, map_mode=args.map_mode)
        ms = r.time_device(steps, 3, barrier)
        stage, counters, pair_counts, tr, bin_off = r.profile(3)
        e2e = r.time_host(steps, 2, barrier)
        roof, _, mean_it = roofline_of(r, stage, pair_counts, tr, bin_off, sm_max_mhz, peak_src)
        n = data.n_correspondences
        rows.append({"config": f"C{cfg}", "workload": f"BASELINE configs[{cfg - 1}] ({NAMES[cfg]}), {data.n_frame_pairs} frame pair(s) per step"
                     + (" as one growth series" if series else ""), "segment_pairs": int(data.n_segments), "correspondences": int(n),
                     "ms_per_step": ms, "value": n / (ms / 1e3), "e2e_ms_per_step": e2e, "e2e_value": n / (e2e / 1e3),
                     "dominant_kernel": roof["kernel"], "frac": roof["frac"],
                     "per_kernel_frac": {k: round(v["frac"], 4) for k, v in roof["per_kernel"].items()},
                     "stage_ms": {k: round(v, 4) for k, v in stage.items() if v > 0.001}, "mean_nn_passes": mean_it})
        r.close()
    return rows


def cpu_baseline(args):
    """the oracle port on the box's host cores: all cores (process pool over segment pairs) and one core (serial, how the reference
    runs), on bounded samples of the headline generator"""
    cores = host_cores()
    n_plants = max(8, -(-cores // 12) * 4)                      # ~0.15-0.3 core-seconds per maize segment pair: 15-30 core-seconds
    segs = syn().make_series(args.config, n_frames=2, plants=n_plants, seed_offset=100).as_batch().segments()
    if len(segs) >= cores:
        segs = segs[:len(segs) // cores * cores]                 # no core idles while the others finish
    n_proc = min(cores, len(segs))
    v, dt, n = cpu_throughput(segs, n_proc)
    v1, dt1, n1 = cpu_throughput(segs[:12], 1)
    return {"value": v, "unit": UNIT, "cores": n_proc, "kind": "port", "cpu": cpu_model(),
            "sample": f"{len(segs)} segment pairs (frame pairs of the same generator), {n} correspondences, {dt:.1f} s wall on {n_proc} processes "
                      "(oracle port: sklearn KDTree + numpy SVD + per-point blend loop, as the reference runs)",
            "ref_1core": {"value": v1, "cores": 1, "sample": f"12 segment pairs (one frame pair), {n1} correspondences, {dt1:.1f} s, serial in one process (how the reference actually runs)"}}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
