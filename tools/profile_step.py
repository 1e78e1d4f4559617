"""one device-resident step of the headline workload (or --config N --frame-pairs F independent pairs), batch in one piece, for ncu:
   LIP_DEVICE_CHUNKS=1 ncu --metrics gpu__time_duration.sum ... python tools/profile_step.py [steps]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
native = importlib.import_module("4d_plant_analysis_b200._native")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
data = syn.make_series(2, n_frames=17, plants=4)
ctx = native.Context(0)
ctx.set_map_dtype(np.int32)
ctx.plan(data.off1, data.off2, data.lm_off)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (data.frames, data.lm1, data.lm2)]
T21 = torch.empty(int(data.off2[-1]), dtype=torch.int32, device=dev); T12 = torch.empty(int(data.off1[-1]), dtype=torch.int32, device=dev)
st = torch.cuda.Stream(dev)
for _ in range(steps):
    ctx.run_series_device(d[0].data_ptr(), len(data.frames), data.row1, data.row2, d[1].data_ptr(), d[2].data_ptr(), T21.data_ptr(), T12.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
print("ok", ctx.counters())
