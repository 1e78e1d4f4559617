"""map mode 2 (final map on the tensor cores) against map mode 0 on a few batches: identical maps, counters, stage times"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
L = importlib.import_module("4d_plant_analysis_b200")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
which = sys.argv[1:] or ["c1s", "tiny", "c3s", "c2s", "c5s", "c2"]
cases = {"c1s": lambda: L.make_config(1, scale=0.2), "c3s": lambda: L.make_config(3, scale=0.3), "c2s": lambda: L.make_config(2, scale=0.34),
         "c5s": lambda: L.make_config(5, scale=0.05), "c2": lambda: L.make_config(2, frame_pairs=16), "c1": lambda: L.make_config(1, frame_pairs=16)}
def tiny():
    rng = np.random.default_rng(5)
    segs = []
    for n1, n2, K in [(5, 7, 2), (40, 3, 2), (12, 12, 3), (1, 9, 1), (30, 31, 1), (64, 90, 4), (300, 1, 1), (1, 300, 1), (129, 257, 2)]:
        Lx = 3.5 * K
        mk = lambda n: np.stack([rng.uniform(0, Lx, n), rng.normal(0, .8, n), rng.normal(0, .8, n)], 1) + 20.0
        lm = np.stack([np.linspace(0.5, Lx - 0.5, K) if K > 1 else np.array([Lx / 2]), np.zeros(K), np.zeros(K)], 1) + 20.0
        segs.append((mk(n1), mk(n2) + 0.2, lm, lm + 0.2))
    return L.batch_from_segments(segs)
cases["tiny"] = tiny
for name in which:
    b = cases[name]()
    out = {}
    for mode in (0, 2):
        ctx = L.Context(0)
        ctx.set_map_mode(mode)
        ctx.plan(b.off1, b.off2, b.lm_off)
        T = ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2)
        t = time.perf_counter()
        for _ in range(3):
            ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2)
        dt = (time.perf_counter() - t) / 3
        ctx.set_profiling(True)
        ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2)
        st = ctx.stage_ms()
        out[mode] = (T, dt, st["final_map_kernel"], ctx.counters()["map_f64_rescans"])
        ctx.close()
    d21 = int((out[0][0][0] != out[2][0][0]).sum()); d12 = int((out[0][0][1] != out[2][0][1]).sum())
    print(f"{name}: S={b.n_segments} corr={b.n_correspondences}  map diffs T21 {d21} T12 {d12} | final map kernel {out[0][2]:.3f} -> {out[2][2]:.3f} ms | "
          f"f64 rescans {out[0][3]} -> {out[2][3]} | host call {1e3 * out[0][1]:.2f} -> {1e3 * out[2][1]:.2f} ms", flush=True)
