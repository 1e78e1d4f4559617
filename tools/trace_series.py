"""LIP_TRACE_PIPELINE=1 python tools/trace_series.py [config n_frames plants] : timeline of the sub-batches of one end-to-end series step (pinned buffers)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
native = importlib.import_module("4d_plant_analysis_b200._native")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
cfg, nf, pl = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 17, 4)
data = syn.make_series(cfg, n_frames=nf, plants=pl)
ctx = native.Context(0)
ctx.set_map_dtype(np.int32)
print(f"{data.n_correspondences} correspondences, {len(data.frames)} rows, {len(data.off1) - 1} segment pairs", file=sys.stderr)
ctx.plan(data.off1, data.off2, data.lm_off)
pins = [torch.from_numpy(a).pin_memory() for a in (data.frames, data.lm1, data.lm2)]
h21 = torch.empty(int(data.off2[-1]), dtype=torch.int32).pin_memory(); h12 = torch.empty(int(data.off1[-1]), dtype=torch.int32).pin_memory()
for i in range(4):
    if i == 3: print("---- traced step ----", file=sys.stderr, flush=True)
    t = time.perf_counter()
    ctx.run_series_ptr(pins[0].data_ptr(), len(data.frames), data.row1, data.row2, pins[1].data_ptr(), pins[2].data_ptr(), h21.data_ptr(), h12.data_ptr())
    print(f"step {i}: {1e3 * (time.perf_counter() - t):.3f} ms", file=sys.stderr, flush=True)
