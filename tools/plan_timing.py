import importlib, os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
L = importlib.import_module("4d_plant_analysis_b200")
ctx = L.default_context()
for cfg, fp in ((2, 64), (1, 64), (5, 2), (1, 1)):
    b = L.make_config(cfg, frame_pairs=fp)
    b2 = L.make_config(cfg, frame_pairs=max(1, fp - 1)) if fp > 1 else L.make_config(3, frame_pairs=1)
    best = 1e9
    for _ in range(6):
        ctx.plan(b2.off1, b2.off2, b2.lm_off)
        t = time.perf_counter(); ctx.plan(b.off1, b.off2, b.lm_off); best = min(best, time.perf_counter() - t)
    print(f"plan of config {cfg} x {fp} ({b.n_segments} segment pairs): {1e3 * best:.3f} ms")
