"""duration of ONE heavy ICP (n_src x n_tgt, K == 1) under ncu, to compare icp_big_kernel and icp_cluster_kernel:
   LIP_CLUSTER_SIZE=8 ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/cluster_item_probe.py 979 1058"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
L = importlib.import_module("4d_plant_analysis_b200")
n_src, n_tgt = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5)
def tube(n, Lx, r=1.5, jitter=0.05):
    s = rng.uniform(0, Lx, n); th = rng.uniform(0, 2 * np.pi, n); rr = r + rng.normal(0, jitter, n)
    return np.stack([s + rng.normal(0, jitter, n), rr * np.cos(th) + 0.02 * s * s / Lx, rr * np.sin(th)], 1)
tgt = tube(n_tgt, 30.0) + 50.0
c, s_ = np.cos(0.05), np.sin(0.05)
src = (tube(n_src, 30.0) + 50.0) @ np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]]).T + [0.4, -0.3, 0.2]
ctx = L.default_context()
for _ in range(3):
    R, t, it, st = ctx.icp(src, [0, n_src], tgt, [0, n_tgt])
print("passes", it, "status", st)
