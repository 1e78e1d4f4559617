"""one device-resident step of `--config N --frame-pairs F` independent frame pairs (batch in one piece), for ncu launch lists:
   LIP_DEVICE_CHUNKS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_config.py 3 16 [steps]
   without ncu it prints the stage times of the profiling mode and the counters"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
native = importlib.import_module("4d_plant_analysis_b200._native")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
cfg, fp = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
data = syn.make_config(cfg, frame_pairs=fp)
ctx = native.Context(0)
ctx.plan(data.off1, data.off2, data.lm_off)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (data.xyz1, data.xyz2, data.lm1, data.lm2)]
T21 = torch.empty(int(data.off2[-1]), dtype=torch.int64, device=dev); T12 = torch.empty(int(data.off1[-1]), dtype=torch.int64, device=dev)
st = torch.cuda.Stream(dev)
def step():
    ctx.run_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), T21.data_ptr(), T12.data_ptr(), st.cuda_stream)
for _ in range(steps):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record(st)
    for _ in range(5):
        step()
    e1.record(st)
torch.cuda.synchronize()
print("ms/step %.3f" % (e0.elapsed_time(e1) / 5), ctx.counters())
if os.environ.get("LIP_PROFILE_COUNTERS"):
    ctx.set_profiling(True)
    step(); torch.cuda.synchronize()
    print("stage_ms", {k: round(v, 4) for k, v in ctx.stage_ms().items() if v > 0.002}, ctx.counters())
    ctx.set_profiling(False)
