"""host-to-host latency of the drop-in calls the unmodified reference driver makes per matched segment pair and direction
(p2p_matching_seg_icp.py:173-176): get_p2p_mapping_from_segments_using_icp on tomato-sized segment pairs, a new shape every call."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
L = importlib.import_module("4d_plant_analysis_b200")
native = importlib.import_module("4d_plant_analysis_b200._native")


class Mesh:
    def __init__(self, v): self.vertices, self.vertex_colors = v, None


b = L.make_config(1)                      # one tomato frame pair: 30 segment pairs
segs = b.segments()
meshes = [(Mesh(x1), Mesh(x2), l1, l2) for x1, x2, l1, l2 in segs]
for m1, m2, l1, l2 in meshes[:3]:
    L.get_p2p_mapping_from_segments_using_icp(m1, m2, l1, l2, 0.4)
for rep in range(3):
    t = time.perf_counter()
    for m1, m2, l1, l2 in meshes:
        L.get_p2p_mapping_from_segments_using_icp(m1, m2, l1, l2, 0.4)
        L.get_p2p_mapping_from_segments_using_icp(m2, m1, l2, l1, 0.4)
    dt = time.perf_counter() - t
    print(f"drop-in get_p2p_mapping_from_segments_using_icp: {1e6 * dt / (2 * len(meshes)):.1f} us per call (60 calls, {b.n_correspondences} correspondences)")
ctx = native.Context(0)
for rep in range(3):
    t = time.perf_counter()
    for x1, x2, l1, l2 in segs:
        ctx.plan([0, len(x1)], [0, len(x2)], [0, len(l1)], native.DIR_T21)
        ctx.run_host(x1, x2, l1, l2)
        ctx.plan([0, len(x2)], [0, len(x1)], [0, len(l1)], native.DIR_T21)
        ctx.run_host(x2, x1, l2, l1)
    dt = time.perf_counter() - t
    print(f"Context.plan + run_host (ctypes):                {1e6 * dt / (2 * len(segs)):.1f} us per call")
tp = tr = 0.0
for x1, x2, l1, l2 in segs:
    t = time.perf_counter(); ctx.plan([0, len(x1)], [0, len(x2)], [0, len(l1)], native.DIR_T21); tp += time.perf_counter() - t
    t = time.perf_counter(); ctx.run_host(x1, x2, l1, l2); tr += time.perf_counter() - t
print(f"  of which plan {1e6 * tp / len(segs):.1f} us, run {1e6 * tr / len(segs):.1f} us")
# same shape again and again (plan fast path)
x1, x2, l1, l2 = segs[0]
ctx.plan([0, len(x1)], [0, len(x2)], [0, len(l1)], native.DIR_T21)
t = time.perf_counter()
for _ in range(200):
    ctx.plan([0, len(x1)], [0, len(x2)], [0, len(l1)], native.DIR_T21)
    ctx.run_host(x1, x2, l1, l2)
print(f"same shape repeated: {1e6 * (time.perf_counter() - t) / 200:.1f} us per call")
t = time.perf_counter()
for _ in range(20):
    L.local_icp_batch(b)
print(f"the whole frame pair as ONE batched call: {1e6 * (time.perf_counter() - t) / 20:.1f} us ({1e6 * (time.perf_counter() - t) / 20 / 60:.1f} us per map)")
# where does the batched call spend its time?
def tm(label, fn, n=20):
    fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    print(f"  {label}: {1e6 * (time.perf_counter() - t) / n:.1f} us")
c2 = native.Context(0)
c2.plan(b.off1, b.off2, b.lm_off)
tm("plan (same shape)", lambda: c2.plan(b.off1, b.off2, b.lm_off))
tm("run_host, validation off", lambda: c2.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2))
c2.set_input_validation(True)
tm("run_host, validation on", lambda: c2.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2))
c2.set_input_validation(False)
tm("local_icp_batch(ctx=c2, validate=False)", lambda: L.local_icp_batch(b, ctx=c2, validate=False))
tm("local_icp_batch(ctx=c2)", lambda: L.local_icp_batch(b, ctx=c2))
tm("local_icp_batch()", lambda: L.local_icp_batch(b))
c2.set_profiling(True)
c2.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2)
print("  stage ms:", {k: round(v, 4) for k, v in c2.stage_ms().items()})
