"""Short device-resident run of the hot path for ncu captures: `python tools/profile_step.py [config] [frame_pairs] [steps]`."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
native = importlib.import_module("4d_plant_analysis_b200._native")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b = syn.make_config(cfg, frame_pairs=fp)
ctx = native.Context(0)
ctx.plan(b.off1, b.off2, b.lm_off)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (b.xyz1, b.xyz2, b.lm1, b.lm2)]
T21 = torch.empty(int(b.off2[-1]), dtype=torch.int64, device=dev)
T12 = torch.empty(int(b.off1[-1]), dtype=torch.int64, device=dev)
for _ in range(steps):
    ctx.run_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), T21.data_ptr(), T12.data_ptr(), 0)
torch.cuda.synchronize()
print("ok", b.n_correspondences, ctx.counters())
