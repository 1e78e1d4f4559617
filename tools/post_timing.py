"""Timing of the whole per-segment body on the GPU: hot path -> double_direction_refinement -> evaluate_p2p_matching
   python tools/post_timing.py [config] [frame_pairs]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("4d_plant_analysis_b200")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
b = L.make_config(cfg, frame_pairs=fp)
ctx = L.default_context()
def timed(f, n=3):
    f(); best = 1e9
    for _ in range(n):
        t = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t)
    return r, best
res, t_map = timed(lambda: L.local_icp_batch(b, validate=False))
_, t_all = timed(lambda: L.match_segments_batch(b))
T21r, t_ref = timed(lambda: L.refine_batch(b, res.T21, res.T12))
sc, t_ev = timed(lambda: L.evaluate_batch(b, T21r, res.T12, voxel_size=0.8, sample_proportion=0.1))
xi, t_int = timed(lambda: L.interpolate_batch(b, T21r, res.T12, 0.5))
print(f"config {cfg} x {fp}: {b.n_segments} segment pairs, {b.n_correspondences} correspondences (host numpy in -> numpy out)")
print(f"  hot path (local_icp_batch)        {1e3 * t_map:8.2f} ms")
print(f"  double_direction_refinement       {1e3 * t_ref:8.2f} ms   changed {100 * (T21r != res.T21).mean():.1f}% of T21")
print(f"  all three in one call (match_segments_batch) {1e3 * t_all:8.2f} ms")
print(f"  evaluate_p2p_matching             {1e3 * t_ev:8.2f} ms   mean scores {sc.mean(0).round(4)}")
print(f"  interpolation_segment_level       {1e3 * t_int:8.2f} ms   {xi.nbytes / 1e6:.0f} MB out, {(b.xyz1.nbytes + b.xyz2.nbytes + 16 * len(xi)) / 1e6:.0f} MB in (PCIe-bound)")
