"""Developer diagnostics on a GPU box: per-case parity statistics against the golden vectors and
rough timings.  Writes gpurun_out/debug.txt.  Not part of the product or the tests."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
L = importlib.import_module("4d_plant_analysis_b200")
out = open(os.path.join(ROOT, "gpurun_out", "debug.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n"); out.flush()
ctx = L.default_context()
for case in ["tiny", "c1_small", "c2_small", "c3_small", "c5_small"]:
    g = load_golden(case)
    b = L.SegmentPairBatch(g["xyz1"], g["xyz2"], g["off1"], g["off2"], g["lm1"], g["lm2"], g["lm_off"], np.zeros(len(g["off1"])-1, np.int32))
    res = L.local_icp_batch(b, details=True)
    raw = res.transforms["raw"]
    P("==", case, "counters", res.counters)
    for d in ("21", "12"):
        v = g["v"+d]; st = raw["status"+d]
        P("  dir", d, "valid mismatch", int(((st != 0) != v).sum()), "degenerate", int((st == 2).sum()), "of", int(v.sum()))
        ok = v & (st == 1)
        itm = raw["iters"+d][ok] != g["it"+d][ok]
        dR = np.abs(raw["R"+d][ok] - g["R"+d][ok]).reshape(ok.sum(), -1).max(1) if ok.any() else np.zeros(0)
        dt = np.abs(raw["t"+d][ok] - g["t"+d][ok]).max(1) if ok.any() else np.zeros(0)
        P("    iters mismatch", int(itm.sum()), "max dR", dR.max() if len(dR) else 0, "max dt", dt.max() if len(dt) else 0,
          "n(dR>1e-5)", int((dR > 1e-5).sum()), "n(dt>1e-5)", int((dt > 1e-5).sum()))
        if itm.any():
            idx = np.nonzero(ok)[0][itm]
            P("    iter mismatch rows", idx[:10], raw["iters"+d][idx][:10], g["it"+d][idx][:10])
        T = res.T21 if d == "21" else res.T12
        P("    map mismatches", int((T != g["T"+d]).sum()), "of", len(T))
for cfg, scale in [(1, 1.0), (2, 1.0), (3, 1.0), (5, 1.0)]:
    b = L.make_config(cfg, scale=scale)
    ctx.plan(b.off1, b.off2, b.lm_off)
    for rep in range(3):
        t = time.perf_counter(); T21, T12 = ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2); dt = time.perf_counter() - t
    P(f"C{cfg}: S={b.n_segments} corr={b.n_correspondences} host e2e {dt*1e3:.3f} ms -> {b.n_correspondences/dt/1e6:.2f} Mcorr/s", ctx.counters())
out.close()
