"""Where the host-to-host time of the numpy API goes: plan / run with pageable numpy / run with pinned buffers.
   python tools/host_path_timing.py [config] [frame_pairs]"""
import ctypes, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("4d_plant_analysis_b200")
native = importlib.import_module("4d_plant_analysis_b200._native")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
b = L.make_config(cfg, frame_pairs=fp)
ctx = L.default_context()
def best(f, n=5):
    f(); t = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); f(); t = min(t, time.perf_counter() - t0)
    return 1e3 * t
print(f"config {cfg} x {fp}: {b.n_correspondences} correspondences, {(b.xyz1.nbytes + b.xyz2.nbytes) / 1e6:.0f} MB in, {8 * b.n_correspondences / 1e6:.0f} MB out")
print(f"  plan                         {best(lambda: ctx.plan(b.off1, b.off2, b.lm_off)):8.2f} ms")
T21, T12 = np.empty(int(b.off2[-1]), np.int64), np.empty(int(b.off1[-1]), np.int64)
print(f"  run_host, pageable numpy     {best(lambda: ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2, T21, T12)):8.2f} ms   (LIP_STAGING_THREADS={os.environ.get('LIP_STAGING_THREADS', 'default')})")
ref21, ref12 = T21.copy(), T12.copy()
lib = native.load()
def pinned(a):
    p = ctypes.c_void_p()
    assert lib.lip_host_alloc(ctypes.byref(p), a.nbytes) == 0
    ctypes.memmove(p, a.ctypes.data, a.nbytes)
    return p.value
px1, px2, pl1, pl2 = (pinned(np.ascontiguousarray(a)) for a in (b.xyz1, b.xyz2, b.lm1, b.lm2))
o21, o12 = pinned(T21), pinned(T12)
print(f"  run_host, pinned buffers     {best(lambda: ctx.run_host_ptr(px1, px2, pl1, pl2, o21, o12)):8.2f} ms")
import ctypes as _c
got21 = np.frombuffer((_c.c_int64 * len(T21)).from_address(o21), dtype=np.int64); got12 = np.frombuffer((_c.c_int64 * len(T12)).from_address(o12), dtype=np.int64)
assert (got21 == ref21).all() and (got12 == ref12).all(), "pageable and pinned runs differ"
print(f"  local_icp_batch(validate=False) {best(lambda: L.local_icp_batch(b, validate=False)):8.2f} ms")
print(f"  local_icp_batch(validate=True)  {best(lambda: L.local_icp_batch(b)):8.2f} ms")
print(f"  memcpy of the inputs (1 thread) {best(lambda: (b.xyz1.copy(), b.xyz2.copy())):8.2f} ms")
